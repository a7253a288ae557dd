// Attention kernels (CUDA cores; head dim 64).  All softmax arithmetic in fp32.
//   enc_attention       : 16 baseline queries vs. a static pyramid view (Tk in {1024,256,64,4}), flash-style
//                         single pass with online softmax; one block per (image, head), 4 queries per warp.
//   dec_self_attention  : one new query per (row, head) against the KV cache through the beam-ancestry table.
//   dec_cross_attention : one query per (row, head) against the 16 memory tokens of the row's image.
#include "kernels.cuh"

namespace fpnmt {

#define LAUNCH_CHECK() FPNMT_CUDA_OK(cudaGetLastError())
constexpr int DH = 64;

// Load 64 contiguous channels [col, col+64) of one row into 64 floats spread as 2 per lane (lane*2, lane*2+1).
__device__ __forceinline__ float2 ld_pair(const Act& a, size_t row, int col, int lane) {
  const bf16* q = a.p + row * (size_t)a.ld + col + lane * 2;
  __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(q);
  float2 v = __bfloat1622float2(h);
  if (a.lo) {
    float2 l = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(q + a.lo));
    v.x += l.x;
    v.y += l.y;
  }
  return v;
}
__device__ __forceinline__ void st_pair(const Act& a, size_t row, int col, int lane, float x, float y) {
  bf16* q = a.p + row * (size_t)a.ld + col + lane * 2;
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  *reinterpret_cast<__nv_bfloat162*>(q) = h;
  if (a.lo) {
    float2 hf = __bfloat1622float2(h);
    *reinterpret_cast<__nv_bfloat162*>(q + a.lo) = __floats2bfloat162_rn(x - hf.x, y - hf.y);
  }
}
// Full 64-wide row segment into registers of ONE lane (used when a lane owns a key).
__device__ __forceinline__ void ld_row64(const Act& a, size_t row, int col, float* f) {
#pragma unroll
  for (int i = 0; i < 8; ++i) ld_act8(a, row, col + i * 8, f + i * 8);
}

// ---------------------------------------------------------------------------------------- encoder
// block = 128 threads = 4 warps; warp w handles queries 4w..4w+3 of the 16; lanes own keys in chunks of 32
// for the score pass and own 2 output dims for the value pass.
__global__ void __launch_bounds__(128) k_enc_attention(Act q, int q_col, Act kv, int k_col, int v_col, int Tq, int Tk,
                                                       Act out, int out_col) {
  pdl_launch();
  pdl_wait();
  __shared__ float sq[16][DH];
  const int b = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale = 0.125f;   // 1/sqrt(64)
  for (int i = threadIdx.x; i < Tq * DH; i += blockDim.x) {
    const int qi = i / DH, d = i % DH;
    sq[qi][d] = ld_act(q, (size_t)b * Tq + qi, q_col + h * DH + d) * scale;
  }
  __syncthreads();
  constexpr int QW = 4;
  float m[QW], l[QW], o0[QW], o1[QW];
#pragma unroll
  for (int i = 0; i < QW; ++i) { m[i] = -INFINITY; l[i] = 0.f; o0[i] = 0.f; o1[i] = 0.f; }
  for (int k0 = 0; k0 < Tk; k0 += 32) {
    const int key = k0 + lane;
    float sc[QW];
#pragma unroll
    for (int i = 0; i < QW; ++i) sc[i] = -INFINITY;
    if (key < Tk) {
      float kr[DH];
      ld_row64(kv, (size_t)b * Tk + key, k_col + h * DH, kr);
#pragma unroll
      for (int i = 0; i < QW; ++i) {
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) a = fmaf(sq[warp * QW + i][d], kr[d], a);
        sc[i] = a;
      }
    }
    float p[QW];
#pragma unroll
    for (int i = 0; i < QW; ++i) {
      const float mn = fmaxf(m[i], warp_max(sc[i]));
      const float corr = __expf(m[i] - mn);
      p[i] = (key < Tk) ? __expf(sc[i] - mn) : 0.f;
      l[i] = l[i] * corr + warp_sum(p[i]);
      o0[i] *= corr;
      o1[i] *= corr;
      m[i] = mn;
    }
    const int kmax = min(32, Tk - k0);
    for (int j = 0; j < kmax; ++j) {
      const float2 v = ld_pair(kv, (size_t)b * Tk + k0 + j, v_col + h * DH, lane);
#pragma unroll
      for (int i = 0; i < QW; ++i) {
        const float pj = __shfl_sync(0xffffffffu, p[i], j);
        o0[i] = fmaf(pj, v.x, o0[i]);
        o1[i] = fmaf(pj, v.y, o1[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < QW; ++i) {
    const int qi = warp * QW + i;
    if (qi < Tq) {
      const float inv = 1.f / l[i];
      st_pair(out, (size_t)b * Tq + qi, out_col + h * DH, lane, o0[i] * inv, o1[i] * inv);
    }
  }
}
int launch_enc_attention(Act q, int q_col, Act kv, int k_col, int v_col, int B, int Tq, int Tk, int heads, Act out,
                         int out_col, cudaStream_t s) {
  if (Tq > 16) {
    set_last_error("enc_attention: Tq must be <= 16");
    return 1;
  }
  dim3 grid(B, heads);
  FPNMT_CUDA_OK(launch_k(k_enc_attention, dim3(grid), dim3(128), 0, s, q, q_col, kv, k_col, v_col, Tq, Tk, out, out_col));
  LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------- decoder attention
// Warp-level attention of ONE query over Tk cached positions.  The query (pre-scaled) lives in shared memory
// (broadcast reads).  Score pass: lane j owns position k0+j and reads its 128 B key row with 8 independent 16 B
// loads (memory-level parallelism, few registers); value pass: lanes own 2 output dims and the probabilities are
// broadcast by shuffle, loads unrolled.  `krow(pos)` maps a position to the row of the K/V views.
__device__ __forceinline__ float dot_row64(const Act& a, size_t row, int col, const float* __restrict__ qs) {
  const bf16* p = a.p + row * (size_t)a.ld + col;
  uint4 h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = *reinterpret_cast<const uint4*>(p + i * 8);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float f[8];
    unpack8(h[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(qs[i * 8 + j], f[j], acc);
  }
  if (a.lo) {
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = *reinterpret_cast<const uint4*>(p + a.lo + i * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float f[8];
      unpack8(h[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(qs[i * 8 + j], f[j], acc);
    }
  }
  return acc;
}

template <typename RowFn>
__device__ __forceinline__ void warp_attend(const float* __restrict__ qs, const Act& kc, int k_col, const Act& vc,
                                            int v_col, int Tk, RowFn krow, int lane, float& m, float& l, float& o0,
                                            float& o1) {
  for (int k0 = 0; k0 < Tk; k0 += 32) {
    const int pos = k0 + lane;
    float sc = -INFINITY;
    unsigned myrow = 0;
    if (pos < Tk) {
      myrow = (unsigned)krow(pos);
      sc = dot_row64(kc, myrow, k_col, qs);
    }
    const float mn = fmaxf(m, warp_max(sc));
    const float corr = __expf(m - mn);
    const float p = (pos < Tk) ? __expf(sc - mn) : 0.f;
    l = l * corr + warp_sum(p);
    o0 *= corr;
    o1 *= corr;
    m = mn;
    const int kmax = min(32, Tk - k0);
#pragma unroll 8
    for (int j = 0; j < kmax; ++j) {
      const unsigned r = __shfl_sync(0xffffffffu, myrow, j);
      const float pj = __shfl_sync(0xffffffffu, p, j);
      const float2 v = ld_pair(vc, r, v_col, lane);
      o0 = fmaf(pj, v.x, o0);
      o1 = fmaf(pj, v.y, o1);
    }
  }
}

constexpr int DEC_WARPS = 4;

// one warp per (row, head).  Position t = *step is the new token: its K/V come from `qkv` and are appended to the
// cache at [row][t]; positions t' < t are read from cache row anc[row][t'] (the beam's ancestor at that time).
__global__ void __launch_bounds__(DEC_WARPS * 32) k_dec_self_attention(Act qkv, Act kc, Act vc,
                                                                       const int* __restrict__ anc_base, size_t anc_stride,
                                                                       const int* __restrict__ step, int rows, int T,
                                                                       int heads, Act out) {
  __shared__ float sq[DEC_WARPS][DH];
  pdl_launch();
  pdl_wait();
  const int gw = blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
  if (gw >= rows * heads) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int row = gw / heads, h = gw % heads;
  const int d = heads * DH;
  const int t = *step;
  const int* anc = anc_base + (size_t)(t & 1) * anc_stride + (size_t)row * T;
  const float2 qv = ld_pair(qkv, row, h * DH, lane);
  const float2 kn = ld_pair(qkv, row, d + h * DH, lane);
  const float2 vn = ld_pair(qkv, row, 2 * d + h * DH, lane);
  sq[w][lane * 2] = qv.x * 0.125f;       // 1/sqrt(64)
  sq[w][lane * 2 + 1] = qv.y * 0.125f;
  st_pair(kc, (size_t)row * T + t, h * DH, lane, kn.x, kn.y);     // append the new K/V to the cache
  st_pair(vc, (size_t)row * T + t, h * DH, lane, vn.x, vn.y);
  __syncwarp();
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  warp_attend(sq[w], kc, h * DH, vc, h * DH, t, [&](int pos) { return (size_t)anc[pos] * T + pos; }, lane, m, l, o0, o1);
  {   // the new position itself (K/V still in registers)
    const float sc = warp_sum((qv.x * kn.x + qv.y * kn.y) * 0.125f);
    const float mn = fmaxf(m, sc);
    const float corr = __expf(m - mn);
    const float p = __expf(sc - mn);
    l = l * corr + p;
    o0 = o0 * corr + p * vn.x;
    o1 = o1 * corr + p * vn.y;
  }
  const float inv = 1.f / l;
  st_pair(out, row, h * DH, lane, o0 * inv, o1 * inv);
}
int launch_dec_self_attention(Act qkv, Act kcache, Act vcache, const int* anc, const int* step, int rows, int T,
                              int heads, Act out, cudaStream_t s) {
  const int warps = rows * heads;
  FPNMT_CUDA_OK(launch_k(k_dec_self_attention, dim3((warps + DEC_WARPS - 1) / DEC_WARPS), dim3(DEC_WARPS * 32), 0, s,
                         qkv, kcache, vcache, anc, (size_t)rows * T, step, rows, T, heads, out));
  return 0;
}

__global__ void __launch_bounds__(DEC_WARPS * 32) k_dec_cross_attention(Act q, Act kv, int k_col, int v_col, int rows,
                                                                        int beam, int Tk, int heads, Act out) {
  __shared__ float sq[DEC_WARPS][DH];
  pdl_launch();
  pdl_wait();
  const int gw = blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
  if (gw >= rows * heads) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int row = gw / heads, h = gw % heads;
  const int img = row / beam;
  const float2 qv = ld_pair(q, row, h * DH, lane);
  sq[w][lane * 2] = qv.x * 0.125f;
  sq[w][lane * 2 + 1] = qv.y * 0.125f;
  __syncwarp();
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  warp_attend(sq[w], kv, k_col + h * DH, kv, v_col + h * DH, Tk, [&](int pos) { return (size_t)img * Tk + pos; }, lane,
              m, l, o0, o1);
  const float inv = 1.f / l;
  st_pair(out, row, h * DH, lane, o0 * inv, o1 * inv);
}
int launch_dec_cross_attention(Act q, Act kv, int k_col, int v_col, int rows, int beam, int Tk, int heads, Act out,
                               cudaStream_t s) {
  const int warps = rows * heads;
  FPNMT_CUDA_OK(launch_k(k_dec_cross_attention, dim3((warps + DEC_WARPS - 1) / DEC_WARPS), dim3(DEC_WARPS * 32), 0, s,
                         q, kv, k_col, v_col, rows, beam, Tk, heads, out));
  return 0;
}

}  // namespace fpnmt
