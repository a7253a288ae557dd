// Host-side construction of CUtensorMap descriptors (TMA) and of IgemmOp launch plans.
#pragma once
#include "igemm.cuh"

namespace fpnmt {

// Geometry of one implicit-GEMM problem, in logical (unsplit) terms.
struct ConvGeom {
  int N, H, W;        // output (== input) pixel grid; dense: N=1,H=1,W=rows
  int Cin, Cout;
  int kh, kw;         // filter taps
  int pad_y, pad_x;   // leading zero padding (top / left); trailing padding is implied by OOB fill
};

// Build an IgemmOp.  `in` is the input activation view (C == Cin), `wt` the device weight matrix
// [Cout][kh*kw*Cin] bf16 (split mode: [Cout][2*kh*kw*Cin], low halves after the high halves),
// `split` selects the BF16X3 three-term product.
int make_igemm_op(IgemmOp* op, const ConvGeom& g, const Act& in, const bf16* wt, bool split, const float* bias, int act,
                  const Act& out, float* out_f32, int ld_f32, int res_mode, const Act& res, int num_sms, int force_bn = 0,
                  bool tma_store = true, bool b_stationary = true);

int encode_tmap_act(CUtensorMap* m, const bf16* base, int C, int ld, int W, int H, int N, int tw, int th, int bn);
int encode_tmap_2d(CUtensorMap* m, const bf16* base, uint64_t inner, uint64_t outer, uint64_t ld_elems, int box_outer);

}  // namespace fpnmt
