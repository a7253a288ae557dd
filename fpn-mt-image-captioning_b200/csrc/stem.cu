// Fused stem convolution (see stem.cuh).  Warp roles of the 416-thread CTA:
//   warps 0..3  : epilogue.  tcgen05.ld of the tile's accumulator (lane = output pixel), + bias, activation, bf16 pack,
//                 transpose through a swizzled per-warp staging buffer, 512 contiguous bytes per store instruction.
//   warp 4      : TMEM allocation, weight TMA (once), MMA issue (one thread): KSTEPS tcgen05.mma (M = 128, N = Cout, K = 16)
//                 per tile into accumulator (tile & 1).
//   warps 5..12 : producers.  16-byte cp.async of the fp32 input patch of the NEXT tile while the current one is converted to
//                 a bf16 2x2 space-to-depth patch (12 channels per pixel) and the im2col rows are copied out of it with
//                 8-byte shared loads into the K-major SWIZZLE_128B layout (16-byte unit u of row r at
//                 r*128 + ((u ^ (r & 7)) << 4) inside a 16 KB chunk of 64 K-elements), then fence.proxy.async + arrive.
#include "stem.cuh"

#include "tensormap.cuh"

namespace fpnmt {

constexpr int ST_TW = 64, ST_TH = 2;
constexpr int ST_PROD = 256;                       // producer threads
constexpr int ST_THREADS = 160 + ST_PROD;          // 4 epilogue warps + 1 MMA warp + 8 producer warps
constexpr int ST_A_CHUNK = 128 * 128;              // 128 pixels x 64 K-elements, bf16

// The stride-2 KH x KH convolution on 3 channels is evaluated as a stride-1 KS x KS convolution on the 2x2 space-to-depth
// image (12 channels, KS = ceil(KH / 2)); taps that fall outside the original window carry zero weights.  K index:
//   k = ky' * (KS*12) + kx' * 12 + dy * 6 + dx * 3 + c      <->   original tap (2ky'+dy-SH, 2kx'+dx-SH, c)
// so the K-elements of one (pixel, ky') are KS*12 CONTIGUOUS bf16 of a space-to-depth row: the im2col tile is built with
// 8-byte shared-memory loads instead of one 4-byte load per element.
template <int KH, int PAD, int COUT>
struct StemCfg {
  static constexpr int KS = (KH + 1) / 2;                              // taps of the space-to-depth convolution
  static constexpr int PADS = (PAD + 1) / 2;                           // its padding
  static constexpr int SH = 2 * PADS - PAD;                            // original tap = 2*tap' + d - SH
  static constexpr int SEG = KS * 12;                                  // K-elements per (pixel, ky')
  static constexpr int KTOT = KS * SEG;
  static constexpr int KSTEPS = KTOT / 16;
  static constexpr int UNITS = KTOT / 8;                               // 16-byte units per im2col row
  static constexpr int SEGU = SEG / 8;
  static constexpr int KCH = (UNITS + 7) / 8;                          // 64-element chunks
  static constexpr int PROWS = 2 * (ST_TH + KS - 1);                   // fp32 patch: input rows
  static constexpr int PV = (2 * (ST_TW + KS - 1) * 3 + 3) / 4;        //             float4 per row
  static constexpr int PROWF = PV * 4;
  static constexpr int SROWS = ST_TH + KS - 1;                         // space-to-depth patch rows
  static constexpr int SPX = ST_TW + KS - 1;                           //                      pixels per row (24 B each)
  static constexpr int SPITCH = 68;                                    // row pitch in pixels: 6*68 = 24 (mod 32) words keeps
                                                                       // the 8-byte gathers of the im2col copy conflict-free
  static constexpr int W_CHUNK = COUT * 128;
  static constexpr int OFF_A = KCH * W_CHUNK;
  static constexpr int OFF_IN = OFF_A + 2 * KCH * ST_A_CHUNK;
  static constexpr int OFF_S2D = OFF_IN + 2 * PROWS * PROWF * 4;
  static constexpr int OFF_OUT = (OFF_S2D + SROWS * SPITCH * 24 + 15) / 16 * 16;
  static constexpr int OFF_BIAS = OFF_OUT + 4 * 32 * COUT * 2;
  static constexpr int OFF_BARS = OFF_BIAS + COUT * 4;
  static constexpr int SMEM = OFF_BARS + 128 + 1024;
  static constexpr int TMEM_COLS = 2 * COUT < 32 ? 32 : 2 * COUT;
  static_assert(W_CHUNK % 1024 == 0 && OFF_IN % 16 == 0 && OFF_S2D % 16 == 0 && OFF_BARS % 8 == 0, "alignment");
  static_assert(SPITCH >= SPX && KTOT % 16 == 0 && SEG % 8 == 0 && COUT % 32 == 0 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "shape");
};

__device__ __forceinline__ void st_cp16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ uint32_t st_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int KH, int PAD, int COUT>
__global__ void __launch_bounds__(ST_THREADS, 1) stem_kernel(const __grid_constant__ CUtensorMap tmW, const StemParams p) {
  using C = StemCfg<KH, PAD, COUT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;
  uint8_t* sA = smem + C::OFF_A;
  float* sIn = reinterpret_cast<float*>(smem + C::OFF_IN);
  uint8_t* sS = smem + C::OFF_S2D;
  uint8_t* sOut = smem + C::OFF_OUT;
  float* sBias = reinterpret_cast<float*>(smem + C::OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BARS);
  uint64_t* fullW = bars;
  uint64_t* fullA = bars + 1;     // [2] im2col tile written (256 arrivals)
  uint64_t* emptyA = bars + 3;    // [2] MMAs of the tile have read it
  uint64_t* tfull = bars + 5;     // [2] accumulator complete
  uint64_t* tempty = bars + 7;    // [2] accumulator drained (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef FPNMT_DBG_STAMPS   // build.py --dbg-stamps (see igemm.cu): timeline of tile 10 (CTA 0)
  long long* dbg = (p.dbg && blockIdx.x == 0) ? p.dbg + 16 : nullptr;
#define SDBG(k) do { if (dbg && it == 10) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); dbg[k] = t_; } } while (0)
#else
#define SDBG(k) do { } while (0)
#endif
  pdl_launch();
  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmW);
      mbar_init(fullW, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&fullA[i], ST_PROD);
        mbar_init(&emptyA[i], 1);
        mbar_init(&tfull[i], 1);
        mbar_init(&tempty[i], 128);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<C::TMEM_COLS>(tmem_slot);
  }
  for (int i = threadIdx.x; i < COUT; i += ST_THREADS) sBias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ntl = ((int)blockIdx.x < p.tiles) ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 4) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    // convergent warp, one elected lane issues (see umma_bf16_pred in common.cuh)
    const uint32_t leader = elect_one() ? 1u : 0u;
    {
      mbar_expect_tx_pred(fullW, C::KCH * C::W_CHUNK, leader);
      for (int c = 0; c < C::KCH; ++c) tma_load_2d_pred(sW + c * C::W_CHUNK, &tmW, fullW, c * 64, 0, leader);
      mbar_wait(fullW, 0);
      constexpr uint32_t IDESC = umma_idesc_bf16(128, COUT);
      const uint64_t w0 = umma_desc_sw128(smem_u32(sW));
      for (int it = 0; it < ntl; ++it) {
        const int b = it & 1, ph = (it >> 1) & 1;
        mbar_wait(&tempty[b], ph ^ 1);
        mbar_wait(&fullA[b], ph);
        tc_fence_after();
        if (leader) SDBG(7);
        const uint64_t a0 = umma_desc_sw128(smem_u32(sA + b * C::KCH * ST_A_CHUNK));
#pragma unroll
        for (int ks = 0; ks < C::KSTEPS; ++ks) {
          const int c = ks >> 2, k = ks & 3;
          umma_bf16_pred(tmem_base + b * COUT, a0 + (uint64_t)(c * (ST_A_CHUNK >> 4) + 2 * k),
                    w0 + (uint64_t)(c * (C::W_CHUNK >> 4) + 2 * k), IDESC, ks > 0 ? 1u : 0u, leader);
        }
        umma_commit_pred(&emptyA[b], leader);
        umma_commit_pred(&tfull[b], leader);
        if (leader) SDBG(8);
      }
    }
    __syncwarp();
  } else if (warp > 4) {
    // ------------------------------------------------------------------------------------------ producers
    const int ptid = threadIdx.x - 160;
    pdl_wait();
    const float* __restrict__ img = *p.img_slot;
    const int rowlen = p.W * 3;
    auto stage = [&](int it) {                                  // fp32 patch of tile `it`: 16-byte cp.async, zero fill
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, n = t / (p.tiles_x * p.tiles_y);
      const int y_in0 = 2 * (ty * ST_TH - C::PADS);
      const int fc0 = 6 * (tx * ST_TW - C::PADS);                // first float column (multiple of 4)
      const uint32_t dst0 = smem_u32(sIn + (it & 1) * C::PROWS * C::PROWF);
      constexpr int NV = (C::PROWS * C::PV + ST_PROD - 1) / ST_PROD;
#pragma unroll
      for (int j = 0; j < NV; ++j) {                             // fully unrolled: independent address chains
        const int v = ptid + j * ST_PROD;
        if (v < C::PROWS * C::PV) {
          const int ry = v / C::PV, vj = v - ry * C::PV;
          const int y = y_in0 + ry, fc = fc0 + 4 * vj;
          const bool ok = y >= 0 && y < p.H && fc >= 0 && fc + 3 < rowlen;
          const float* src = ok ? img + ((size_t)n * p.H + y) * rowlen + fc : img;
          st_cp16(dst0 + (ry * C::PROWF + vj * 4) * 4, src, ok);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (ntl > 0) stage(0);
    for (int it = 0; it < ntl; ++it) {
      const int b = it & 1, ph = (it >> 1) & 1;
      if (ptid == 0) {
        SDBG(0);
#ifdef FPNMT_DBG_STAMPS
        if (dbg && it == 11) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); dbg[12] = t_; }
#endif
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync 2, 256;" ::: "memory");           // patch `it` complete; everyone is done with tile it-1
      if (ptid == 0) SDBG(1);
      if (it + 1 < ntl) stage(it + 1);
      if (ptid == 0) SDBG(2);
      // ---- fp32 patch -> bf16 space-to-depth patch: one (y', x') pixel = 2 rows x 6 floats -> 12 bf16 (24 bytes)
      const float* __restrict__ sp = sIn + b * C::PROWS * C::PROWF;
      for (int i = ptid; i < C::SROWS * C::SPX; i += ST_PROD) {
        const int sy = i / C::SPX, sx = i - sy * C::SPX;
        const float2* r0 = reinterpret_cast<const float2*>(sp + (2 * sy) * C::PROWF + 6 * sx);
        const float2* r1 = reinterpret_cast<const float2*>(sp + (2 * sy + 1) * C::PROWF + 6 * sx);
        const float2 a0 = r0[0], a1 = r0[1], a2 = r0[2], b0 = r1[0], b1 = r1[1], b2 = r1[2];
        uint2* d = reinterpret_cast<uint2*>(sS + (size_t)(sy * C::SPITCH + sx) * 24);
        d[0] = make_uint2(st_pack(a0.x, a0.y), st_pack(a1.x, a1.y));
        d[1] = make_uint2(st_pack(a2.x, a2.y), st_pack(b0.x, b0.y));
        d[2] = make_uint2(st_pack(b1.x, b1.y), st_pack(b2.x, b2.y));
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");           // space-to-depth patch complete
      if (ptid == 0) SDBG(3);
      mbar_wait(&emptyA[b], ph ^ 1);
      if (ptid == 0) SDBG(4);
      uint8_t* ab = sA + b * C::KCH * ST_A_CHUNK;
      if constexpr (C::UNITS % 8 == 0) {
        // every chunk has 8 units: a thread keeps ONE unit column (ptid & 7) and walks 4 rows per chunk; all loads of the
        // tile are independent and issued back to back
        const int ul = ptid & 7, r0 = ptid >> 3;
        const int sw = (ul ^ (r0 & 7)) << 4;                     // (row & 7) == (r0 & 7) for row = r0 + 32 j
        uint2 lo[C::KCH][4], hi[C::KCH][4];
#pragma unroll
        for (int c = 0; c < C::KCH; ++c) {
          const int u = c * 8 + ul, kyp = u / C::SEGU, part = u - kyp * C::SEGU;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int row = r0 + 32 * j;
            const uint2* src = reinterpret_cast<const uint2*>(sS + ((size_t)((row >> 6) + kyp) * C::SPITCH + (row & 63)) * 24 + part * 16);
            lo[c][j] = src[0];
            hi[c][j] = src[1];
          }
        }
#pragma unroll
        for (int c = 0; c < C::KCH; ++c)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(ab + c * ST_A_CHUNK + (r0 + 32 * j) * 128 + sw) =
                make_uint4(lo[c][j].x, lo[c][j].y, hi[c][j].x, hi[c][j].y);
      } else {
        static_assert(C::KCH == 1, "partial chunks only for single-chunk stems");
        constexpr int NI = (128 * C::UNITS + ST_PROD - 1) / ST_PROD;
        uint2 lo[NI], hi[NI];
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int idx = ptid + j * ST_PROD;
          const int row = idx / C::UNITS, u = idx - row * C::UNITS;
          const int kyp = u / C::SEGU, part = u - kyp * C::SEGU;
          const uint2* src = reinterpret_cast<const uint2*>(sS + ((size_t)((row >> 6) + kyp) * C::SPITCH + (row & 63)) * 24 + part * 16);
          if (idx < 128 * C::UNITS) {
            lo[j] = src[0];
            hi[j] = src[1];
          }
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int idx = ptid + j * ST_PROD;
          const int row = idx / C::UNITS, u = idx - row * C::UNITS;
          if (idx < 128 * C::UNITS)
            *reinterpret_cast<uint4*>(ab + row * 128 + ((u ^ (row & 7)) << 4)) = make_uint4(lo[j].x, lo[j].y, hi[j].x, hi[j].y);
        }
      }
      if (ptid == 0) SDBG(5);
      fence_proxy_async();                                     // generic-proxy writes -> visible to the tensor core
      mbar_arrive(&fullA[b]);
      if (ptid == 0) SDBG(6);
    }
  } else {
    // ------------------------------------------------------------------------------------------ epilogue
    constexpr int UPR = COUT / 8;                              // 16-byte units per output pixel
    uint8_t* so = sOut + warp * 32 * COUT * 2;
    // activation as a branch-free clamp: none (-inf, inf), ReLU (0, inf), ReLU6 (0, 6)
    const float act_lo = p.act == ACT_NONE ? -INFINITY : 0.f, act_hi = p.act == ACT_RELU6 ? 6.f : INFINITY;
    pdl_wait();                                                // the output buffer may still be read by an earlier kernel
    for (int it = 0; it < ntl; ++it) {
      const int b = it & 1, ph = (it >> 1) & 1;
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, n = t / (p.tiles_x * p.tiles_y);
      mbar_wait(&tfull[b], ph);
      tc_fence_after();
      if (threadIdx.x == 0) SDBG(9);
      const int swz = (UPR == 8) ? (lane & 7) : ((lane >> 1) & (UPR - 1));
#pragma unroll
      for (int h = 0; h < COUT / 32; ++h) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + b * COUT + h * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 b0 = *reinterpret_cast<const float4*>(sBias + h * 32 + u * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(sBias + h * 32 + u * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x0 = fminf(fmaxf(__uint_as_float(r[u * 8 + i * 2]) + bb[i * 2], act_lo), act_hi);
            const float x1 = fminf(fmaxf(__uint_as_float(r[u * 8 + i * 2 + 1]) + bb[i * 2 + 1], act_lo), act_hi);
            w[i] = st_pack(x0, x1);
          }
          *reinterpret_cast<uint4*>(so + lane * (COUT * 2) + (((h * 4 + u) ^ swz) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[b]);
      __syncwarp();
      if (threadIdx.x == 0) SDBG(10);
      // warp w holds tile rows 32w .. 32w+31 = 32 consecutive pixels of output row ty*2 + (w >> 1)
      const size_t pix0 = ((size_t)n * p.Ho + ty * ST_TH + (warp >> 1)) * p.Wo + tx * ST_TW + (warp & 1) * 32;
      uint8_t* g = reinterpret_cast<uint8_t*>(p.out.p + pix0 * (size_t)p.out.ld);
#pragma unroll
      for (int i = 0; i < UPR; ++i) {
        const int q = i * 32 + lane, row = q / UPR, u = q % UPR;
        const int sw = (UPR == 8) ? (row & 7) : ((row >> 1) & (UPR - 1));
        const uint4 v = *reinterpret_cast<const uint4*>(so + row * (COUT * 2) + ((u ^ sw) << 4));
        *reinterpret_cast<uint4*>(g + (size_t)q * 16) = v;
      }
      __syncwarp();
      if (threadIdx.x == 0) SDBG(11);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// Permute the folded weights [Cout][Kp] (k = (ky*KH + kx)*3 + c) into the space-to-depth K order [Cout][KTOT].
__global__ void k_stem_permute(const bf16* __restrict__ w, int Kp, int cout, int kh, int ks, int sh, bf16* __restrict__ wp) {
  const int seg = ks * 12, ktot = ks * seg;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * ktot) return;
  const int co = i / ktot, k = i % ktot;
  const int kyp = k / seg, r = k % seg, kxp = r / 12, e = r % 12;
  const int ky = 2 * kyp + e / 6 - sh, kx = 2 * kxp + (e % 6) / 3 - sh, c = e % 3;
  const bool ok = ky >= 0 && ky < kh && kx >= 0 && kx < kh;
  wp[i] = ok ? w[(size_t)co * Kp + (ky * kh + kx) * 3 + c] : __float2bfloat16_rn(0.f);
}

// ------------------------------------------------------------------------------------------------ host
int stem_set_attributes() {
  FPNMT_CUDA_OK(cudaFuncSetAttribute(stem_kernel<7, 3, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, StemCfg<7, 3, 64>::SMEM));
  FPNMT_CUDA_OK(cudaFuncSetAttribute(stem_kernel<3, 0, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, StemCfg<3, 0, 32>::SMEM));
  return 0;
}

int make_stem_op(StemOp* op, const float* const* img_slot, int N, int H, int W, int kh, int pad, int cout, const bf16* wt, int Kp,
                 bf16* wt_s2d, const float* bias, int act, const Act& out, int num_sms) {
  const bool resnet_like = kh == 7 && pad == 3 && cout == 64, mobilenet = kh == 3 && pad == 0 && cout == 32;
  if (!resnet_like && !mobilenet) {
    set_last_error("make_stem_op: only the 7x7/pad 3/64 and 3x3/pad 0/32 stems are instantiated");
    return 1;
  }
  if (act != ACT_NONE && act != ACT_RELU && act != ACT_RELU6) {
    set_last_error("make_stem_op: activation must be none, ReLU or ReLU6");
    return 1;
  }
  if (W % 128 || H % 4 || out.lo || out.ld != cout || (Kp % 8)) {
    set_last_error("make_stem_op: image width must be a multiple of 128, height of 4, plain bf16 output with ld == Cout");
    return 1;
  }
  StemParams& p = op->p;
  p = StemParams{};
  p.img_slot = img_slot;
  p.N = N; p.H = H; p.W = W; p.Ho = H / 2; p.Wo = W / 2;
  p.tiles_x = p.Wo / ST_TW;
  p.tiles_y = p.Ho / ST_TH;
  p.tiles = N * p.tiles_x * p.tiles_y;
  p.bias = bias;
  p.act = act;
  p.out = out;
  op->kh = kh;
  op->cout = cout;
  op->grid = p.tiles < num_sms ? p.tiles : num_sms;
  op->flops = 2.0 * (double)N * p.Ho * p.Wo * (kh * kh * 3) * cout;
  const int ks = (kh + 1) / 2, sh = 2 * ((pad + 1) / 2) - pad, ktot = ks * ks * 12;
  k_stem_permute<<<(cout * ktot + 255) / 256, 256>>>(wt, Kp, cout, kh, ks, sh, wt_s2d);
  FPNMT_CUDA_OK(cudaGetLastError());
  FPNMT_CUDA_OK(cudaDeviceSynchronize());
  return encode_tmap_2d(&op->tmW, wt_s2d, (uint64_t)ktot, (uint64_t)cout, (uint64_t)ktot, cout);
}

int stem_launch(const StemOp& op, cudaStream_t stream) {
  if (op.kh == 7)
    FPNMT_CUDA_OK(launch_k(stem_kernel<7, 3, 64>, dim3(op.grid), dim3(ST_THREADS), (size_t)StemCfg<7, 3, 64>::SMEM, stream, op.tmW, op.p));
  else
    FPNMT_CUDA_OK(launch_k(stem_kernel<3, 0, 32>, dim3(op.grid), dim3(ST_THREADS), (size_t)StemCfg<3, 0, 32>::SMEM, stream, op.tmW, op.p));
  return 0;
}

}  // namespace fpnmt
