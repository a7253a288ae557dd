// Implicit-GEMM convolution / dense kernel for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM
// -> fused epilogue.  One kernel serves
//   * Dense layers and 1x1 convolutions   (A = [rows, K] activations, 1 tap)
//   * kxk stride-1 "same"/explicit-pad convolutions on NHWC maps (A tiles are spatial boxes fetched per
//     filter tap with the tap offset added to the TMA coordinates; out-of-bounds = zero padding)
// with out[pixel, co] = act( sum_{tap,ci} A[pixel+tap, ci] * Wt[co, tap*Cin+ci] + bias[co] + residual ).
#pragma once
#include "common.cuh"

namespace fpnmt {

constexpr int IG_BM = 128;   // rows (pixels) per CTA tile == UMMA_M == TMEM lanes
constexpr int IG_BK = 64;    // K elements per pipeline stage == one 128 B swizzle row of bf16
constexpr int IG_THREADS = 320;   // TMA warp + MMA warp + 8 epilogue warps

enum ResMode { RES_NONE = 0, RES_SAME = 1, RES_UP2 = 2 };

struct IgemmParams {
  // output pixel space (N images of H x W); a dense layer is N=1,H=1,W=rows
  int N, H, W;
  int tw, th, bn;                 // M-tile box (tw*th*bn == 128)
  int tiles_x, tiles_y, tiles_n;  // number of boxes along W, H, N
  int tiles_co;                   // ceil(Cout / BN)
  int taps_y, taps_x, pad_y, pad_x;
  int Cin, kchunks;               // kchunks = ceil(Cin / 64)
  int nterms;                     // 1 = bf16, 3 = bf16x3 split (A.hi*W.hi + A.hi*W.lo + A.lo*W.hi)
  int b_lo_off;                   // K offset of the low halves of the weights (split mode)
  int Cout;
  const float* bias;              // [Cout] or nullptr
  int act;                        // ActFn
  Act out;                        // bf16 output view (p may be nullptr)
  float* out_f32;                 // optional fp32 output [pixel][ld_f32]
  int ld_f32;
  int res_mode;                   // ResMode; residual is added before the activation
  Act res;                        // RES_SAME: same pixel grid; RES_UP2: (N, H/2, W/2) nearest-upsampled
  long long* dbg;                 // optional timeline buffer (globaltimer stamps of block 0), normally nullptr
  int b_stat;                     // 1: the whole weight panel [BN][K] stays in shared memory for every tile of the CTA (one output-
                                  // channel tile, bf16 mode, panel <= 96 KB): the ring then carries only the activation chunks
  int tma_store;                  // 1: bf16 outputs leave through bulk tensor stores (output map in the tmA_lo slot; bf16 mode,
                                  // BN >= 128): see the epilogue
};

struct IgemmOp {
  CUtensorMap tmA_hi, tmA_lo, tmB;
  IgemmParams p;
  int BN;     // 32 / 64 / 128 / 256
  int smem_bytes;   // dynamic shared memory of the launch
  int grid;
  double flops;   // algorithmic FLOPs (2*MAC, one term) for reporting
};

int igemm_stages(int BN);
size_t igemm_smem_bytes(int BN, int b_chunks = 0);   // b_chunks > 0: stationary weight panel of that many 64-wide k-chunks
int igemm_launch(const IgemmOp& op, cudaStream_t stream);
int igemm_set_attributes();   // cudaFuncSetAttribute for every instantiation (call once per device)

}  // namespace fpnmt
