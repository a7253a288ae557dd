// Bandwidth-bound NHWC kernels: 128-bit loads/stores over 8 channels per thread, grids sized to the data.
#include "kernels.cuh"

namespace fpnmt {

static inline int nblocks(size_t work, int threads) { return (int)((work + threads - 1) / threads); }
#define LAUNCH_CHECK() FPNMT_CUDA_OK(cudaGetLastError())

// ---------------------------------------------------------------------------------------- im2col (stem)
__global__ void k_im2col_stem(const float* const* __restrict__ img_slot, int N, int H, int W, int kh, int kw, int stride, int pad_t,
                              int pad_l, int Ho, int Wo, Act out) {
  pdl_launch();
  pdl_wait();
  const float* __restrict__ img = *img_slot;
  const int groups = out.C / 8;
  const size_t total = (size_t)N * Ho * Wo * groups;
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const size_t pix = idx / groups;
  const int xo = (int)(pix % Wo);
  const int yo = (int)((pix / Wo) % Ho);
  const int n = (int)(pix / ((size_t)Wo * Ho));
  const int ktot = kh * kw * 3;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = g * 8 + i;
    float val = 0.f;
    if (k < ktot) {
      const int c = k % 3;
      const int kx = (k / 3) % kw;
      const int ky = k / (3 * kw);
      const int y = yo * stride + ky - pad_t;
      const int x = xo * stride + kx - pad_l;
      if (y >= 0 && y < H && x >= 0 && x < W) val = __ldg(img + (((size_t)n * H + y) * W + x) * 3 + c);
    }
    v[i] = val;
  }
  st_act8(out, pix, g * 8, v);
}
// Tiled stride-2 variant: a block stages the input rows/columns feeding a 64 x 4 tile of output pixels in shared memory
// with coalesced loads (each image element leaves L2 about once instead of once per tap), then writes the im2col rows
// (64 * Kpad contiguous bf16 per output row) with 16-byte stores.
template <int KH, int KW, int PAD>
__global__ void __launch_bounds__(256) k_im2col_stem_s2(const float* const* __restrict__ img_slot, int N, int H, int W, int Ho,
                                                        int Wo, Act out) {
  constexpr int TW = 64, TH = 4;
  constexpr int ROWS = (TH - 1) * 2 + KH;
  constexpr int ROWF = ((TW - 1) * 2 + KW) * 3;
  constexpr int KTOT = KH * KW * 3;
  __shared__ float s[ROWS][ROWF];
  pdl_launch();
  pdl_wait();
  const float* __restrict__ img = *img_slot;
  const int xo0 = blockIdx.x * TW, yo0 = blockIdx.y * TH, n = blockIdx.z;
  const int x_in0 = xo0 * 2 - PAD, y_in0 = yo0 * 2 - PAD;
  for (int i = threadIdx.x; i < ROWS * ROWF; i += blockDim.x) {
    const int ry = i / ROWF, j = i % ROWF;
    const int x = x_in0 + j / 3, y = y_in0 + ry;
    float v = 0.f;
    if (x >= 0 && x < W && y >= 0 && y < H) v = __ldg(img + (((size_t)n * H + y) * W + x) * 3 + (j % 3));
    s[ry][j] = v;
  }
  __syncthreads();
  const int groups = out.C / 8;
  for (int w = threadIdx.x; w < TH * TW * groups; w += blockDim.x) {
    const int g = w % groups, pl = w / groups;
    const int xl = pl % TW, yl = pl / TW;
    if (xo0 + xl >= Wo || yo0 + yl >= Ho) continue;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = g * 8 + i;
      v[i] = (k < KTOT) ? s[yl * 2 + k / (KW * 3)][xl * 6 + k % (KW * 3)] : 0.f;
    }
    st_act8(out, ((size_t)n * Ho + yo0 + yl) * Wo + xo0 + xl, g * 8, v);
  }
}

int launch_im2col_stem(const float* const* img, int N, int H, int W, int kh, int kw, int stride, int pad_t, int pad_l, int Ho,
                       int Wo, Act out, cudaStream_t s) {
  if (stride == 2 && pad_t == pad_l && Ho <= 65535 && N <= 65535) {
    const dim3 grid((Wo + 63) / 64, (Ho + 3) / 4, N);
    if (kh == 7 && kw == 7 && pad_t == 3) {
      FPNMT_CUDA_OK(launch_k(k_im2col_stem_s2<7, 7, 3>, grid, dim3(256), 0, s, img, N, H, W, Ho, Wo, out));
      return 0;
    }
    if (kh == 3 && kw == 3 && pad_t == 0) {
      FPNMT_CUDA_OK(launch_k(k_im2col_stem_s2<3, 3, 0>, grid, dim3(256), 0, s, img, N, H, W, Ho, Wo, out));
      return 0;
    }
  }
  const size_t total = (size_t)N * Ho * Wo * (out.C / 8);
  FPNMT_CUDA_OK(launch_k(k_im2col_stem, dim3(nblocks(total, 256)), dim3(256), 0, s, img, N, H, W, kh, kw, stride, pad_t, pad_l, Ho, Wo, out));
  LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------- pooling
// K = compile-time window (2 or 3; 0 = generic run-time window): the fixed-size variants are fully unrolled so that all
// window loads are issued back to back.
template <int K>
__global__ void k_maxpool(Act in, int N, int H, int W, int k_rt, int stride, int pad_t, int pad_l, int Ho, int Wo,
                          int zero_pad, Act out) {
  pdl_launch();
  pdl_wait();
  const int k = K ? K : k_rt;
  const int groups = in.C / 8;
  const size_t total = (size_t)N * Ho * Wo * groups;
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const size_t pix = idx / groups;
  const int xo = (int)(pix % Wo);
  const int yo = (int)((pix / Wo) % Ho);
  const int n = (int)(pix / ((size_t)Wo * Ho));
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
  if constexpr (K != 0) {
    uint4 raw[K * K];
    bool ok[K * K];
#pragma unroll
    for (int ky = 0; ky < K; ++ky)
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int y = yo * stride + ky - pad_t, x = xo * stride + kx - pad_l;
        ok[ky * K + kx] = y >= 0 && y < H && x >= 0 && x < W;
        raw[ky * K + kx] = ok[ky * K + kx] ? *reinterpret_cast<const uint4*>(in.p + (((size_t)n * H + y) * W + x) * in.ld + g * 8)
                                          : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
    for (int t = 0; t < K * K; ++t) {
      if (ok[t]) {
        float v[8];
        unpack8(raw[t], v);
        if (in.lo) {
          // split activations: add the low halves (rare path: BF16X3 mode)
          float lo[8];
          const int ky = t / K, kx = t % K;
          const int y = yo * stride + ky - pad_t, x = xo * stride + kx - pad_l;
          unpack8(*reinterpret_cast<const uint4*>(in.p + in.lo + (((size_t)n * H + y) * W + x) * in.ld + g * 8), lo);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] += lo[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
      } else if (zero_pad) {
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], 0.f);
      }
    }
  } else {
    for (int ky = 0; ky < k; ++ky) {
      const int y = yo * stride + ky - pad_t;
      for (int kx = 0; kx < k; ++kx) {
        const int x = xo * stride + kx - pad_l;
        if (y < 0 || y >= H || x < 0 || x >= W) {
          if (zero_pad) {
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], 0.f);
          }
          continue;
        }
        float v[8];
        ld_act8(in, ((size_t)n * H + y) * W + x, g * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
      }
    }
  }
  st_act8(out, pix, g * 8, m);
}
int launch_maxpool(Act in, int N, int H, int W, int k, int stride, int pad_t, int pad_l, int Ho, int Wo, bool zero_pad,
                   Act out, cudaStream_t s) {
  const size_t total = (size_t)N * Ho * Wo * (in.C / 8);
  const dim3 grid(nblocks(total, 256));
  if (k == 3)
    FPNMT_CUDA_OK(launch_k(k_maxpool<3>, grid, dim3(256), 0, s, in, N, H, W, k, stride, pad_t, pad_l, Ho, Wo, zero_pad ? 1 : 0, out));
  else if (k == 2)
    FPNMT_CUDA_OK(launch_k(k_maxpool<2>, grid, dim3(256), 0, s, in, N, H, W, k, stride, pad_t, pad_l, Ho, Wo, zero_pad ? 1 : 0, out));
  else
    FPNMT_CUDA_OK(launch_k(k_maxpool<0>, grid, dim3(256), 0, s, in, N, H, W, k, stride, pad_t, pad_l, Ho, Wo, zero_pad ? 1 : 0, out));
  LAUNCH_CHECK();
  return 0;
}

__global__ void k_avgpool2(Act in, int N, int H, int W, Act out) {
  pdl_launch();
  pdl_wait();
  const int groups = in.C / 8;
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * groups;
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const size_t pix = idx / groups;
  const int xo = (int)(pix % Wo);
  const int yo = (int)((pix / Wo) % Ho);
  const int n = (int)(pix / ((size_t)Wo * Ho));
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int dy = 0; dy < 2; ++dy)
    for (int dx = 0; dx < 2; ++dx) {
      float v[8];
      ld_act8(in, ((size_t)n * H + 2 * yo + dy) * W + 2 * xo + dx, g * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] += v[i];
    }
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] *= 0.25f;
  st_act8(out, pix, g * 8, a);
}
int launch_avgpool2(Act in, int N, int H, int W, Act out, cudaStream_t s) {
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (in.C / 8);
  FPNMT_CUDA_OK(launch_k(k_avgpool2, dim3(nblocks(total, 256)), dim3(256), 0, s, in, N, H, W, out));
  LAUNCH_CHECK();
  return 0;
}

__global__ void k_subsample2(Act in, int N, int H, int W, Act out) {
  pdl_launch();
  pdl_wait();
  const int groups = in.C / 8;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const size_t total = (size_t)N * Ho * Wo * groups;
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const size_t pix = idx / groups;
  const int xo = (int)(pix % Wo);
  const int yo = (int)((pix / Wo) % Ho);
  const int n = (int)(pix / ((size_t)Wo * Ho));
  const size_t ipix = ((size_t)n * H + 2 * yo) * W + 2 * xo;
  const bf16* src = in.p + ipix * (size_t)in.ld + g * 8;
  bf16* dst = out.p + pix * (size_t)out.ld + g * 8;
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);           // exact copy of hi
  if (in.lo) *reinterpret_cast<uint4*>(dst + out.lo) = *reinterpret_cast<const uint4*>(src + in.lo);   // and lo
}
int launch_subsample2(Act in, int N, int H, int W, Act out, cudaStream_t s) {
  const size_t total = (size_t)N * ((H + 1) / 2) * ((W + 1) / 2) * (in.C / 8);
  FPNMT_CUDA_OK(launch_k(k_subsample2, dim3(nblocks(total, 256)), dim3(256), 0, s, in, N, H, W, out));
  LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------- depthwise 3x3
__global__ void k_depthwise3x3(Act in, int N, int H, int W, int stride, int pad_t, int pad_l, int Ho, int Wo,
                               const float* __restrict__ w, const float* __restrict__ bias, int act, Act out) {
  pdl_launch();
  pdl_wait();
  const int C = in.C;
  const int groups = C / 8;
  const size_t total = (size_t)N * Ho * Wo * groups;
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const size_t pix = idx / groups;
  const int xo = (int)(pix % Wo);
  const int yo = (int)((pix / Wo) % Ho);
  const int n = (int)(pix / ((size_t)Wo * Ho));
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = bias ? bias[g * 8 + i] : 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int y = yo * stride + ky - pad_t;
    if (y < 0 || y >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int x = xo * stride + kx - pad_l;
      if (x < 0 || x >= W) continue;
      float v[8];
      ld_act8(in, ((size_t)n * H + y) * W + x, g * 8, v);
      const float4 w0 = *reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + g * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + g * 8 + 4);
      a[0] = fmaf(v[0], w0.x, a[0]); a[1] = fmaf(v[1], w0.y, a[1]); a[2] = fmaf(v[2], w0.z, a[2]); a[3] = fmaf(v[3], w0.w, a[3]);
      a[4] = fmaf(v[4], w1.x, a[4]); a[5] = fmaf(v[5], w1.y, a[5]); a[6] = fmaf(v[6], w1.z, a[6]); a[7] = fmaf(v[7], w1.w, a[7]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = apply_act(a[i], act);
  st_act8(out, pix, g * 8, a);
}
// Stride-1 variant with register tiling along x: one thread produces 4 consecutive output pixels of one 8-channel group,
// so every input value is loaded once per row for (up to) three outputs and the 9x8 weights are fetched once per 4 outputs
// (the one-output-per-thread kernel above issues 9 input + 18 weight loads per output and ran at ~1.1 TB/s).
__global__ void __launch_bounds__(128) k_depthwise3x3_s1x4(Act in, int N, int H, int W, int pad_t, int pad_l, int Ho, int Wo,
                                                            const float* __restrict__ w, const float* __restrict__ bias, int act,
                                                            Act out) {
  pdl_launch();
  pdl_wait();
  const int C = in.C;
  const int groups = C / 8;
  const int Wq = (Wo + 3) / 4;
  const size_t total = (size_t)N * Ho * Wq * groups;
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const size_t q = idx / groups;
  const int xq = (int)(q % Wq);
  const int yo = (int)((q / Wq) % Ho);
  const int n = (int)(q / ((size_t)Wq * Ho));
  const int xo0 = xq * 4;
  float a[4][8];
  {
    float b8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b8[i] = 0.f;
    if (bias) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + g * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + g * 8 + 4));
      b8[0] = b0.x; b8[1] = b0.y; b8[2] = b0.z; b8[3] = b0.w; b8[4] = b1.x; b8[5] = b1.y; b8[6] = b1.z; b8[7] = b1.w;
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[o][i] = b8[i];
  }
  // Branch-free window: coordinates are clamped and out-of-image values zeroed AFTER the load, so all 18 loads of the
  // thread are independent of any control flow and can be in flight together (with a branch per tap the loads of a thread
  // were serialised and the kernel ran at 1.06 TB/s whatever the layer).
  uint4 raw[3][6];
  bool ok[3][6];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int y = yo + ky - pad_t;
    const bool yok = y >= 0 && y < H;
    const int yc = min(max(y, 0), H - 1);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int x = xo0 + j - pad_l;
      ok[ky][j] = yok && x >= 0 && x < W;
      const int xc = min(max(x, 0), W - 1);
      raw[ky][j] = *reinterpret_cast<const uint4*>(in.p + (((size_t)n * H + yc) * W + xc) * in.ld + g * 8);
    }
  }
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    float v[6][8];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      unpack8(raw[ky][j], v[j]);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[j][i] = ok[ky][j] ? v[j][i] : 0.f;
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + g * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + g * 8 + 4));
      const float wk[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[o][i] = fmaf(v[o + kx][i], wk[i], a[o][i]);
    }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    if (xo0 + o < Wo) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[o][i] = apply_act(a[o][i], act);
      st_act8(out, ((size_t)n * Ho + yo) * Wo + xo0 + o, g * 8, a[o]);
    }
  }
}
int launch_depthwise3x3(Act in, int N, int H, int W, int stride, int pad_t, int pad_l, int Ho, int Wo, const float* w,
                        const float* bias, int act, Act out, cudaStream_t s) {
  if (stride == 1 && !in.lo) {
    const size_t total = (size_t)N * Ho * ((Wo + 3) / 4) * (in.C / 8);
    FPNMT_CUDA_OK(launch_k(k_depthwise3x3_s1x4, dim3(nblocks(total, 128)), dim3(128), 0, s, in, N, H, W, pad_t, pad_l, Ho, Wo, w, bias, act, out));
    LAUNCH_CHECK();
    return 0;
  }
  const size_t total = (size_t)N * Ho * Wo * (in.C / 8);
  FPNMT_CUDA_OK(launch_k(k_depthwise3x3, dim3(nblocks(total, 256)), dim3(256), 0, s, in, N, H, W, stride, pad_t, pad_l, Ho, Wo, w, bias, act, out));
  LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------- BN+ReLU (pre-activation)
// A thread keeps ONE 8-channel group (its scale/shift live in 16 registers) and walks pixels with a fixed stride, four
// independent 16-byte loads in flight per iteration.  (The one-item-per-thread version re-read 16 scalar parameters per
// 16-byte item: 8x more L1 wavefronts for the parameters than for the data, 2.2 TB/s.)
__global__ void __launch_bounds__(256) k_scale_shift_relu(Act in, size_t pixels, size_t pix_slots, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, Act out) {
  pdl_launch();
  const int groups = in.C / 8;
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= pix_slots * groups) return;
  const int g = (int)(idx % groups);
  const size_t p0 = idx / groups;
  float sc[8], sh[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(scale + g * 8)), a1 = __ldg(reinterpret_cast<const float4*>(scale + g * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + g * 8)), b1 = __ldg(reinterpret_cast<const float4*>(shift + g * 8 + 4));
    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
  }
  pdl_wait();
  for (size_t p = p0; p < pixels; p += 4 * pix_slots) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t q = p + u * pix_slots;
      if (q < pixels) ld_act8(in, q, g * 8, v[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t q = p + u * pix_slots;
      if (q < pixels) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[u][i] = fmaxf(fmaf(v[u][i], sc[i], sh[i]), 0.f);
        st_act8(out, q, g * 8, v[u]);
      }
    }
  }
}
int launch_scale_shift_relu(Act in, size_t pixels, const float* scale, const float* shift, Act out, cudaStream_t s) {
  const int groups = in.C / 8;
  // about two full waves of 256-thread blocks, each thread looping over its pixels
  size_t slots = ((size_t)148 * 2048 * 2 + groups - 1) / groups;
  if (slots > pixels) slots = pixels;
  if (slots < 1) slots = 1;
  const size_t total = slots * groups;
  FPNMT_CUDA_OK(launch_k(k_scale_shift_relu, dim3(nblocks(total, 256)), dim3(256), 0, s, in, pixels, slots, scale, shift, out));
  LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------- 3x3 convolution to ONE channel
// The co-attention score map (retinanet.py:287, Conv2D(1, 3x3, same) on a 256-channel map) as a CUDA-core kernel: a
// 1-output-channel GEMM wastes 31/32 of the smallest tensor-core tile (the igemm version ran at 5-8 TFLOP/s).  Block = 8x8
// output pixels; the 10x10x256 bf16 halo is staged in shared memory (pixel pitch 576 B: conflict-free 16-byte reads);
// lane = one 8-channel unit whose 9x8 weights live in registers, warp = one row of 8 pixels; each halo value is loaded
// once per row and used by the up to three pixels whose windows contain it; lanes are summed with shuffles.
constexpr int SC_C = 256, SC_PITCH = 576, SC_TILE = 8, SC_HALO = SC_TILE + 2;
__global__ void __launch_bounds__(256) k_conv3x3_c1(Act in, const float* __restrict__ w, const float* __restrict__ bias,
                                                    int N, int H, int W, Act out) {
  extern __shared__ __align__(16) uint8_t s_halo[];
  pdl_launch();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float wr[9][8];                                            // weights of this lane's channel unit (static: before the wait)
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w + t * SC_C + lane * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(w + t * SC_C + lane * 8 + 4));
    wr[t][0] = a.x; wr[t][1] = a.y; wr[t][2] = a.z; wr[t][3] = a.w;
    wr[t][4] = b.x; wr[t][5] = b.y; wr[t][6] = b.z; wr[t][7] = b.w;
  }
  const float bv = __ldg(bias);
  pdl_wait();
  const int x0 = blockIdx.x * SC_TILE, y0 = blockIdx.y * SC_TILE, n = blockIdx.z;
  for (int i = threadIdx.x; i < SC_HALO * SC_HALO * 32; i += 256) {
    const int px = i >> 5, u = i & 31;
    const int hy = px / SC_HALO, hx = px - hy * SC_HALO;
    const int y = y0 + hy - 1, x = x0 + hx - 1;
    const bool ok = y >= 0 && y < H && x >= 0 && x < W;
    const bf16* src = ok ? in.p + (((size_t)n * H + y) * W + x) * in.ld + u * 8 : in.p;
    const int sz = ok ? 16 : 0;                              // src-size 0: zero fill (the "same" padding)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(s_halo + px * SC_PITCH + u * 16)),
                 "l"(src), "r"(sz) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int yl = warp;                                       // output row of the tile
  float acc[SC_TILE];
#pragma unroll
  for (int i = 0; i < SC_TILE; ++i) acc[i] = 0.f;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
    for (int hx = 0; hx < SC_HALO; ++hx) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(s_halo + ((yl + dy) * SC_HALO + hx) * SC_PITCH + lane * 16), f);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xl = hx - dx;                              // output pixel whose tap (dy, dx) reads halo column hx
        if (xl >= 0 && xl < SC_TILE) {
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[xl] = fmaf(f[c], wr[dy * 3 + dx][c], acc[xl]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < SC_TILE; ++i) acc[i] = warp_sum(acc[i]);
  const int y = y0 + yl;
  if (lane < SC_TILE && y < H && x0 + lane < W) {
    float v = acc[0];
#pragma unroll
    for (int i = 1; i < SC_TILE; ++i) v = (lane == i) ? acc[i] : v;
    st_act(out, ((size_t)n * H + y) * W + x0 + lane, 0, v + bv);
  }
}
int launch_conv3x3_c1(Act in, const float* w, const float* bias, int N, int H, int W, Act out, cudaStream_t s) {
  if (in.C != SC_C || in.lo || out.lo) {
    set_last_error("conv3x3_c1: needs a plain bf16 256-channel input");
    return 1;
  }
  const size_t smem = (size_t)SC_HALO * SC_HALO * SC_PITCH;
  const dim3 grid((W + SC_TILE - 1) / SC_TILE, (H + SC_TILE - 1) / SC_TILE, N);
  FPNMT_CUDA_OK(launch_k(k_conv3x3_c1, grid, dim3(256), smem, s, in, w, bias, N, H, W, out));
  return 0;
}

int elementwise_set_attributes() {   // per device (Engine::init)
  FPNMT_CUDA_OK(cudaFuncSetAttribute(k_conv3x3_c1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)SC_HALO * SC_HALO * SC_PITCH)));
  return 0;
}

// ---------------------------------------------------------------------------------------- co-attention
// grid (chunks, N); each block recomputes the image's softmax statistics (HW <= 4096 scores, L2 resident)
// and scales `pix_per_block` pixels x C channels.
__global__ void k_coattention(Act score, Act cls, int HW, int pix_per_block, Act out) {
  pdl_launch();
  pdl_wait();
  __shared__ float red[32];
  __shared__ float s_max, s_inv;
  const int n = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  float m = -INFINITY;
  for (int p = tid; p < HW; p += blockDim.x) m = fmaxf(m, ld_act(score, (size_t)n * HW + p, 0));
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? red[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) s_max = t;
  }
  __syncthreads();
  m = s_max;
  float sum = 0.f;
  for (int p = tid; p < HW; p += blockDim.x) sum += __expf(ld_act(score, (size_t)n * HW + p, 0) - m);
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) s_inv = 1.f / t;
  }
  __syncthreads();
  const float inv = s_inv;
  const int groups = cls.C / 8;
  const int p0 = blockIdx.x * pix_per_block;
  const int work = pix_per_block * groups;
  for (int i = tid; i < work; i += blockDim.x) {
    const int p = p0 + i / groups;
    if (p >= HW) break;
    const int g = i % groups;
    const size_t pix = (size_t)n * HW + p;
    const float wgt = __expf(ld_act(score, pix, 0) - m) * inv;
    float v[8];
    ld_act8(cls, pix, g * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= wgt;
    st_act8(out, pix, g * 8, v);
  }
}
int launch_coattention(Act score, Act cls, int N, int HW, Act out, cudaStream_t s) {
  // every block recomputes the image's softmax statistics (2 passes over HW scores): larger pixel chunks for the large maps
  // keep that overhead below the scaling work itself
  const int ppb = HW >= 4096 ? 256 : HW >= 1024 ? 128 : 32;
  dim3 grid((HW + ppb - 1) / ppb, N);
  FPNMT_CUDA_OK(launch_k(k_coattention, dim3(grid), dim3(256), 0, s, score, cls, HW, ppb, out));
  LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------- LayerNorm family
// one warp per row of C = 512 (16 values per lane, two-pass statistics in registers)
template <typename LoadFn>
__device__ __forceinline__ void ln_row_512(LoadFn load, const float* __restrict__ gamma, const float* __restrict__ beta,
                                           float eps, const float* __restrict__ add, const Act& out, size_t orow,
                                           int lane) {
  float v[16];
  load(lane * 8, v);
  load(256 + lane * 8, v + 8);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.f / 512.f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float d = v[i] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / 512.f) + eps);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = h * 256 + lane * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    if (add) {
      a0 = __ldg(reinterpret_cast<const float4*>(add + c));
      a1 = __ldg(reinterpret_cast<const float4*>(add + c + 4));
    }
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float aa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = ((v[h * 8 + i] - mean) * rstd * gg[i] + bb[i]) + aa[i];
    st_act8(out, orow, c, o);
  }
}

__global__ void k_tokens_ln_pos(Act in, int HW, size_t rows, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, const float* __restrict__ pos, Act out) {
  pdl_launch();
  pdl_wait();
  const size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int p = (int)(row % HW);
  ln_row_512([&](int c, float* f) { ld_act8(in, row, c, f); }, gamma, beta, eps, pos + (size_t)p * 512, out, row, lane);
}
int launch_tokens_ln_pos(Act in, int N, int HW, const float* gamma, const float* beta, float eps, const float* pos,
                         Act out, cudaStream_t s) {
  if (in.C != 512) {
    set_last_error("tokens_ln_pos: d_model must be 512");
    return 1;
  }
  const size_t rows = (size_t)N * HW;
  FPNMT_CUDA_OK(launch_k(k_tokens_ln_pos, dim3(nblocks(rows, 8)), dim3(256), 0, s, in, HW, rows, gamma, beta, eps, pos, out));
  LAUNCH_CHECK();
  return 0;
}

__global__ void k_layernorm_rows(const float* __restrict__ x, int rows, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, Act out) {
  pdl_launch();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (size_t)row * 512;
  ln_row_512(
      [&](int c, float* f) {
        const float4 a = *reinterpret_cast<const float4*>(xr + c);
        const float4 b = *reinterpret_cast<const float4*>(xr + c + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
      },
      gamma, beta, eps, nullptr, out, (size_t)row, lane);
}
int launch_layernorm_rows(const float* x, int rows, int C, const float* gamma, const float* beta, float eps, Act out,
                          cudaStream_t s) {
  if (C != 512) {
    set_last_error("layernorm_rows: d_model must be 512");
    return 1;
  }
  FPNMT_CUDA_OK(launch_k(k_layernorm_rows, dim3(nblocks(rows, 8)), dim3(256), 0, s, x, rows, gamma, beta, eps, out));
  LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------- embedding + position
__global__ void k_embed_pos(const int* __restrict__ tokens, const float* __restrict__ emb, const float* __restrict__ pos,
                            const int* __restrict__ step, int rows, int C, Act out) {
  pdl_launch();
  pdl_wait();
  const int groups = C / 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * groups) return;
  const int g = idx % groups, r = idx / groups;
  const int t = *step;
  const float* e = emb + (size_t)tokens[r] * C + g * 8;
  const float* p = pos + (size_t)t * C + g * 8;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = e[i] + p[i];
  st_act8(out, (size_t)r, g * 8, v);
}
int launch_embed_pos(const int* tokens, const float* emb, const float* pos, const int* step, int rows, int C, Act out,
                     cudaStream_t s) {
  FPNMT_CUDA_OK(launch_k(k_embed_pos, dim3(nblocks((size_t)rows * (C / 8), 256)), dim3(256), 0, s, tokens, emb, pos, step, rows, C, out));
  LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------- input preprocessing
// dataset.py:19-26 after the JPEG decode: tf.image.resize(img, (S,S)) (bilinear, half-pixel centres, no antialias) and
// mobilenet_v2.preprocess_input (x / 127.5 - 1).  uint8 HWC in, float32 NHWC in [-1, 1] out; one thread per output pixel.
__global__ void k_preprocess(const uint8_t* __restrict__ img, int N, int H, int W, int S, float* __restrict__ out) {
  pdl_launch();
  pdl_wait();
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * S * S) return;
  const int x = (int)(idx % S), y = (int)((idx / S) % S), n = (int)(idx / ((size_t)S * S));
  const float sy = (y + 0.5f) * ((float)H / (float)S) - 0.5f;
  const float sx = (x + 0.5f) * ((float)W / (float)S) - 0.5f;
  const float fy0 = floorf(sy), fx0 = floorf(sx);
  const float wy = sy - fy0, wx = sx - fx0;
  const int y0 = min(max((int)fy0, 0), H - 1), y1 = min(max((int)fy0 + 1, 0), H - 1);
  const int x0 = min(max((int)fx0, 0), W - 1), x1 = min(max((int)fx0 + 1, 0), W - 1);
  const uint8_t* base = img + (size_t)n * H * W * 3;
  const uint8_t *p00 = base + ((size_t)y0 * W + x0) * 3, *p01 = base + ((size_t)y0 * W + x1) * 3;
  const uint8_t *p10 = base + ((size_t)y1 * W + x0) * 3, *p11 = base + ((size_t)y1 * W + x1) * 3;
  float* o = out + idx * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float tl = p00[c], tr = p01[c], bl = p10[c], br = p11[c];
    const float top = tl * (1.f - wx) + tr * wx;
    const float bot = bl * (1.f - wx) + br * wx;
    o[c] = (top * (1.f - wy) + bot * wy) / 127.5f - 1.f;
  }
}
int launch_preprocess(const uint8_t* img, int N, int H, int W, int S, float* out, cudaStream_t s) {
  const size_t total = (size_t)N * S * S;
  FPNMT_CUDA_OK(launch_k(k_preprocess, dim3(nblocks(total, 256)), dim3(256), 0, s, img, N, H, W, S, out));
  LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------- conversions
__global__ void k_f32_to_act(const float* __restrict__ x, size_t rows, int C, Act out) {
  pdl_launch();
  pdl_wait();
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= rows * C) return;
  st_act(out, idx / C, (int)(idx % C), x[idx]);
}
int launch_f32_to_act(const float* x, size_t rows, int C, Act out, cudaStream_t s) {
  FPNMT_CUDA_OK(launch_k(k_f32_to_act, dim3(nblocks(rows * C, 256)), dim3(256), 0, s, x, rows, C, out));
  LAUNCH_CHECK();
  return 0;
}
__global__ void k_act_to_f32(Act in, size_t rows, float* __restrict__ out) {
  pdl_launch();
  pdl_wait();
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= rows * in.C) return;
  out[idx] = ld_act(in, idx / in.C, (int)(idx % in.C));
}
int launch_act_to_f32(Act in, size_t rows, float* out, cudaStream_t s) {
  FPNMT_CUDA_OK(launch_k(k_act_to_f32, dim3(nblocks(rows * in.C, 256)), dim3(256), 0, s, in, rows, out));
  LAUNCH_CHECK();
  return 0;
}

}  // namespace fpnmt
