// Fused decoder cross-attention block (see xattn.cuh).  One CTA per image:
//   warp 0     : TMA producer.  Mt_b (128 x 512, 8 chunks) and the first two feature tiles of Nt_b are per-batch constants
//                and are fetched BEFORE griddepcontrol.wait; only the 16 x 512 activation rows depend on the previous kernel.
//   warp 1     : MMA issuer.  chain 1: S^T[128 (h,j) x 16 rows] = Mt_b . out1^T   (32 tcgen05.mma, N = 16)
//                             chain 2: O^T[4 x 128 features x 16 rows] = Nt_b . P^T (4 x 8 tcgen05.mma)
//   warps 2..5 : softmax over the 16 memory tokens (16-lane shuffle groups on the TMEM-loaded scores), P^T -> shared
//                memory (K-major, 128B swizzle, UMMA B operand), then bias + residual + LayerNorm of the output rows.
#include "xattn.cuh"

#include "tensormap.cuh"

namespace fpnmt {

constexpr int XA_CHUNK_A = 128 * 64 * 2;                 // 16 KB: 128 rows x 64 K
constexpr int XA_CHUNK_B = XA_NROWS * 64 * 2;            // 2 KB: 16 rows x 64 K
constexpr int XA_OFF_A1 = 0;                             // 8 chunks (Mt_b); later the LayerNorm scratch
constexpr int XA_OFF_A2 = 8 * XA_CHUNK_A;                // 4 chunks (two feature tiles of Nt_b in flight)
constexpr int XA_OFF_B1 = XA_OFF_A2 + 4 * XA_CHUNK_A;    // 8 chunks (out1 rows)
constexpr int XA_OFF_B2 = XA_OFF_B1 + 8 * XA_CHUNK_B;    // 2 chunks (P^T)
constexpr int XA_OFF_PART = XA_OFF_B2 + 2 * XA_CHUNK_B;  // float2 [8][16]
constexpr int XA_OFF_BARS = XA_OFF_PART + 8 * 16 * 8;
constexpr int XA_SCR_STRIDE = 516;                       // floats per row of the transposed output scratch (16 B aligned
                                                         // rows, 4-bank skew: float4 reads of 8 consecutive rows are conflict-free)
constexpr int XA_TMEM_COLS = 128;

constexpr int XA_OFF_VEC = XA_OFF_BARS + 256;               // float [3][512]: output bias, LayerNorm gamma, beta (static)
size_t xattn_smem_bytes() { return XA_OFF_VEC + 3 * 512 * 4 + 1024; }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(XA_THREADS, 1)
xattn_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmN,
             const __grid_constant__ CUtensorMap tmX, const XattnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA1 = smem + XA_OFF_A1;
  uint8_t* sA2 = smem + XA_OFF_A2;
  uint8_t* sB1 = smem + XA_OFF_B1;
  uint8_t* sB2 = smem + XA_OFF_B2;
  float2* sPart = reinterpret_cast<float2*>(smem + XA_OFF_PART);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + XA_OFF_BARS);
  float* sVec = reinterpret_cast<float*>(smem + XA_OFF_VEC);
  uint64_t* fullA1 = bars;          // Mt_b landed
  uint64_t* fullB1 = bars + 1;      // out1 rows landed
  uint64_t* fullA2 = bars + 2;      // [4] feature tile ft of Nt_b landed
  uint64_t* emptyA2 = bars + 6;     // [2] slot pair free again
  uint64_t* tfull1 = bars + 8;      // scores ready
  uint64_t* pready = bars + 9;      // P^T written (128 arrivals)
  uint64_t* tfull2 = bars + 10;     // outputs ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const int opnd = p.layer * p.Btot + p.b0 + b;                 // which per-image operand set

#ifdef FPNMT_DBG_STAMPS   // build.py --dbg-stamps (see igemm.cu)
  __shared__ long long* s_dbg;
  if (threadIdx.x == 0) {
    s_dbg = nullptr;
    if (p.dbg && blockIdx.x == 0) {
      const long long inst = (long long)atomicAdd((unsigned long long*)p.dbg, 1ull);
      s_dbg = p.dbg + 16 + (inst % 8) * 16;
      long long t_;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
      s_dbg[0] = t_;
    }
  }
#define XDBG(k) do { if (s_dbg) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); s_dbg[k] = t_; } } while (0)
#else
#define XDBG(k) do { } while (0)
#endif
  pdl_launch();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmM);
    tma_prefetch_desc(&tmN);
    tma_prefetch_desc(&tmX);
    mbar_init(fullA1, 1);
    mbar_init(fullB1, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&fullA2[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&emptyA2[i], 1);
    mbar_init(tfull1, 1);
    mbar_init(pready, 128);
    mbar_init(tfull2, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<XA_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, XA_NROWS);

  if (warp == 0) {
    // convergent warp, one elected lane issues (see umma_bf16_pred in common.cuh)
    const uint32_t leader = elect_one() ? 1u : 0u;
    {
      // per-batch constants: before the grid dependency resolves
      mbar_expect_tx_pred(fullA1, 8 * XA_CHUNK_A, leader);
      for (int i = 0; i < 8; ++i) tma_load_2d_pred(sA1 + i * XA_CHUNK_A, &tmM, fullA1, i * 64, opnd * 128, leader);
      for (int ft = 0; ft < 2; ++ft) {
        mbar_expect_tx_pred(&fullA2[ft], 2 * XA_CHUNK_A, leader);
        for (int c = 0; c < 2; ++c)
          tma_load_2d_pred(sA2 + (ft * 2 + c) * XA_CHUNK_A, &tmN, &fullA2[ft], c * 64, opnd * 512 + ft * 128, leader);
      }
      pdl_wait();
      if (leader) XDBG(1);
      mbar_expect_tx_pred(fullB1, 8 * XA_CHUNK_B, leader);
      for (int i = 0; i < 8; ++i) tma_load_2d_pred(sB1 + i * XA_CHUNK_B, &tmX, fullB1, i * 64, b * p.beam, leader);
      // feature tiles 2,3 of Nt_b go into the upper half of the Mt_b area as soon as chain 1 has consumed it, so their
      // load latency overlaps the softmax instead of sitting between the two MMA chains
      mbar_wait(tfull1, 0);
      for (int ft = 2; ft < 4; ++ft) {
        mbar_expect_tx_pred(&fullA2[ft], 2 * XA_CHUNK_A, leader);
        for (int c = 0; c < 2; ++c)
          tma_load_2d_pred(sA1 + (4 + (ft - 2) * 2 + c) * XA_CHUNK_A, &tmN, &fullA2[ft], c * 64, opnd * 512 + ft * 128, leader);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    {
      // ---- chain 1: scores^T
      mbar_wait(fullA1, 0);
      if (leader) XDBG(2);
      mbar_wait(fullB1, 0);
      if (leader) XDBG(3);
      tc_fence_after();
      {
        const uint64_t a0 = umma_desc_sw128(smem_u32(sA1));
        const uint64_t b0 = umma_desc_sw128(smem_u32(sB1));
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_pred(tmem_base, a0 + (uint64_t)(i * (XA_CHUNK_A >> 4) + 2 * k), b0 + (uint64_t)(i * (XA_CHUNK_B >> 4) + 2 * k), IDESC,
                      (i > 0 || k > 0) ? 1u : 0u, leader);
      }
      umma_commit_pred(tfull1, leader);
      if (leader) XDBG(4);
      // ---- chain 2: outputs^T, four feature tiles of 128
      mbar_wait(pready, 0);
      if (leader) XDBG(7);
      tc_fence_after();
      const uint64_t pb0 = umma_desc_sw128(smem_u32(sB2));
      for (int ft = 0; ft < 4; ++ft) {
        mbar_wait(&fullA2[ft], 0);
        if (ft == 2) if (leader) XDBG(8);
        tc_fence_after();
        const uint64_t a0 = umma_desc_sw128(smem_u32(ft < 2 ? sA2 + ft * 2 * XA_CHUNK_A : sA1 + (4 + (ft - 2) * 2) * XA_CHUNK_A));
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_pred(tmem_base + 32 + ft * XA_NROWS, a0 + (uint64_t)(c * (XA_CHUNK_A >> 4) + 2 * k),
                      pb0 + (uint64_t)(c * (XA_CHUNK_B >> 4) + 2 * k), IDESC, (c > 0 || k > 0) ? 1u : 0u, leader);
      }
      umma_commit_pred(tfull2, leader);
      if (leader) XDBG(9);
    }
    __syncwarp();
  } else {
    const int e = warp - 2;                       // 0..3
    const int quarter = warp & 3;                 // TMEM lane window
    const int L = quarter * 32 + lane;            // (head, token) pair == TMEM lane == K index of chain 2
    const int t = e * 32 + lane;                  // 0..127
    const int lrow = t & 15, part = t >> 4;       // LayerNorm phase: row of the image, 64-feature slice
    // static per-layer vectors -> shared memory while the kernel is still waiting for its inputs (128 threads x 3 x 16 B)
    {
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const float* g = v == 0 ? p.obias : v == 1 ? p.gamma : p.beta;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sVec + v * 512 + t * 4)), "l"(g + t * 4) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const float sb = __ldg(p.sbias + (size_t)opnd * XA_PAIRS + L);
    pdl_wait();
    const int grow = b * p.beam + lrow;           // global row handled in the LayerNorm phase
    const bool row_ok = lrow < p.beam && grow < p.R;
    uint4 rres[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) rres[g] = make_uint4(0u, 0u, 0u, 0u);
    if (row_ok) {                                 // residual = out1 row slice: in flight during both MMA chains
      const bf16* q = p.res.p + (size_t)grow * p.res.ld + part * 64;
#pragma unroll
      for (int g = 0; g < 8; ++g) rres[g] = *reinterpret_cast<const uint4*>(q + g * 8);
    }
    // ---- softmax over the 16 tokens of this head, for each of the 16 row columns
    mbar_wait(tfull1, 0);
    tc_fence_after();
    if (t == 0) XDBG(5);
    {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16), r);
      tmem_ld_wait();
      // All 16 columns are processed unconditionally (the padding columns hold finite rows of the next image or zeros)
      // so that the 16 independent shuffle chains interleave; padding columns are zeroed at the end.
      float sc[16], mx[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        sc[c] = __uint_as_float(r[c]) + sb;
        mx[c] = sc[c];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < 16; ++c) mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
      float pr[16], sm[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        pr[c] = __expf(sc[c] - mx[c]);
        sm[c] = pr[c];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < 16; ++c) sm[c] += __shfl_xor_sync(0xffffffffu, sm[c], o);
#pragma unroll
      for (int c = 0; c < 16; ++c) pr[c] = c < p.beam ? __fdividef(pr[c], sm[c]) : 0.f;
      // P^T[row c][k = L]: K-major rows of 128 B, 16-byte units XOR-swizzled with (row & 7)
      uint8_t* base = sB2 + (L >> 6) * XA_CHUNK_B + (L & 7) * 2;
      const int unit = (L & 63) >> 3;
#pragma unroll
      for (int c = 0; c < 16; ++c)
        *reinterpret_cast<bf16*>(base + c * 128 + ((unit ^ (c & 7)) << 4)) = __float2bfloat16_rn(pr[c]);
    }
    fence_proxy_async();                          // generic-proxy writes -> visible to the tensor core's async proxy
    mbar_arrive(pready);
    if (t == 0) XDBG(6);
    // ---- outputs: + bias, transpose through the scratch (aliases the Mt_b area, free since chain 1 completed)
    float* scr = reinterpret_cast<float*>(sA1);
    mbar_wait(tfull2, 0);
    tc_fence_after();
    if (t == 0) XDBG(10);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("bar.sync 1, 128;" ::: "memory");          // every thread's slice of the staged vectors is visible
    {
      uint32_t r[2][32];                          // the four 16-column accumulators are contiguous: two 32-column loads
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + 32, r[0]);
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + 64, r[1]);
      tmem_ld_wait();
#pragma unroll
      for (int ft = 0; ft < 4; ++ft) {
        const int f = ft * 128 + L;
        const float ob = sVec[f];
#pragma unroll
        for (int c = 0; c < 16; ++c) scr[c * XA_SCR_STRIDE + f] = __uint_as_float(r[ft >> 1][(ft & 1) * 16 + c]) + ob;
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (t == 0) XDBG(11);
    // ---- residual + LayerNorm over 512 features: 8 threads per row, 64 features each (Chan's parallel variance)
    float x[64];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 q = *reinterpret_cast<const float4*>(scr + lrow * XA_SCR_STRIDE + part * 64 + i * 4);
      x[4 * i] = q.x; x[4 * i + 1] = q.y; x[4 * i + 2] = q.z; x[4 * i + 3] = q.w;
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float tt[8];
      unpack8(rres[g], tt);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[g * 8 + i] += tt[i];
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) sum += x[i];
    const float mloc = sum * (1.f / 64.f);
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const float d = x[i] - mloc;
      m2 = fmaf(d, d, m2);
    }
    sPart[part * 16 + lrow] = make_float2(mloc, m2);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float mean = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) mean += sPart[q * 16 + lrow].x;
    mean *= 0.125f;
    float M2 = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float2 a = sPart[q * 16 + lrow];
      const float d = a.x - mean;
      M2 += a.y + 64.f * d * d;
    }
    const float rstd = rsqrtf(M2 * (1.f / 512.f) + p.eps);
    if (row_ok) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int f0 = part * 64 + g * 8;
        const float4 ga = *reinterpret_cast<const float4*>(sVec + 512 + f0);
        const float4 gb = *reinterpret_cast<const float4*>(sVec + 512 + f0 + 4);
        const float4 ba = *reinterpret_cast<const float4*>(sVec + 1024 + f0);
        const float4 bb = *reinterpret_cast<const float4*>(sVec + 1024 + f0 + 4);
        const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        const float be[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (x[g * 8 + i] - mean) * rstd * gg[i] + be[i];
        st_act8(p.out, (size_t)grow, f0, o);
      }
    }
  }

  if (threadIdx.x == 64) XDBG(12);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<XA_TMEM_COLS>(tmem_base);
  }
}

int xattn_set_attributes() {
  FPNMT_CUDA_OK(cudaFuncSetAttribute(xattn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xattn_smem_bytes()));
  return 0;
}

int xattn_launch(const XattnOp& op, cudaStream_t stream) {
  FPNMT_CUDA_OK(launch_k(xattn_kernel, dim3(op.p.B), dim3(XA_THREADS), xattn_smem_bytes(), stream, op.tmM, op.tmN, op.tmX, op.p));
  return 0;
}

int make_xattn_op(XattnOp* op, const bf16* Mt, const bf16* Nt, int L, int Btot, int b0, int B, int beam, int layer,
                  const float* sbias, const float* obias, const float* gamma, const float* beta, const Act& x, const Act& out) {
  if (beam > XA_NROWS || x.C != 512 || x.lo || out.lo) {
    set_last_error("make_xattn_op: needs beam <= 16, d_model == 512 and plain bf16 activations");
    return 1;
  }
  XattnParams& p = op->p;
  p = XattnParams{};
  p.B = B;
  p.b0 = b0;
  p.Btot = Btot;
  p.beam = beam;
  p.R = B * beam;
  p.layer = layer;
  p.sbias = sbias;
  p.obias = obias;
  p.gamma = gamma;
  p.beta = beta;
  p.eps = 1e-6f;
  p.res = x;
  p.out = out;
  int rc = encode_tmap_2d(&op->tmM, Mt, 512, (uint64_t)L * Btot * 128, 512, 128);
  if (rc) return rc;
  rc = encode_tmap_2d(&op->tmN, Nt, 128, (uint64_t)L * Btot * 512, 128, 128);
  if (rc) return rc;
  return encode_tmap_2d(&op->tmX, x.p, 512, (uint64_t)B * beam, (uint64_t)x.ld, XA_NROWS);
}

// ------------------------------------------------------------------------------------------------ folding kernels
// grid (B, heads = 8, L), 256 threads.  Keys/values of one (image, head, layer) are staged in shared memory as fp32.
// Fold kernels on mma.sync.m16n8k16 (bf16 in, fp32 accumulate).  The per-(layer, image, head) products have M or N = 16
// (the memory tokens), far below tcgen05's 128-row tiles, and run once per batch (3.2 GFLOP):
//   fold_q:  Mt_b,h[16 j][512 k]  = K_b,h[16 j][64 d] . Wq_h[512 k][64 d]^T / 8      (A = K, B = Wq rows, "col" operand)
//   fold_o:  Nt_b,h[512 f][16 j]  = WoT_h[512 f][64 d] . V_b,h[16 j][64 d]^T          (A = WoT rows, B = V rows)
// Fragments are read straight from global memory in the m16n8k16 register layout (4-byte loads of bf16 pairs); block =
// one (image, head, layer), 4 warps, each warp owns 128 of the 512 k (fold_q) or f (fold_o).
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ld_pair_bf16(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }

__global__ void __launch_bounds__(128) k_xattn_fold_q(Act ckv, int B, int n_mem, const bf16* __restrict__ wq,
                                                      const float* const* __restrict__ bq, bf16* __restrict__ Mt,
                                                      float* __restrict__ sbias) {
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.x, h = blockIdx.y, l = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gr = lane >> 2, gc = (lane & 3) * 2;            // fragment row / column pair of this lane
  const bf16* Kb = ckv.p + (size_t)(b * n_mem) * ckv.ld + l * 1024 + h * 64;
  // A fragments: K_b,h rows j = gr and gr + 8 (zero beyond n_mem), all four k-steps
  uint32_t a[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int d0 = ks * 16 + gc;
    a[ks][0] = gr < n_mem ? ld_pair_bf16(Kb + (size_t)gr * ckv.ld + d0) : 0u;
    a[ks][1] = gr + 8 < n_mem ? ld_pair_bf16(Kb + (size_t)(gr + 8) * ckv.ld + d0) : 0u;
    a[ks][2] = gr < n_mem ? ld_pair_bf16(Kb + (size_t)gr * ckv.ld + d0 + 8) : 0u;
    a[ks][3] = gr + 8 < n_mem ? ld_pair_bf16(Kb + (size_t)(gr + 8) * ckv.ld + d0 + 8) : 0u;
  }
  const bf16* W = wq + (size_t)l * 512 * 512 + h * 64;      // [k][h*64 + d]
  bf16* dst = Mt + ((size_t)(l * B + b) * 128 + h * 16) * 512;
#pragma unroll 4
  for (int nt = 0; nt < 16; ++nt) {
    const int k0 = warp * 128 + nt * 8;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    const bf16* wr = W + (size_t)(k0 + gr) * 512 + gc;      // B fragment: n = k0 + gr, k = d
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma16816(c, a[ks], ld_pair_bf16(wr + ks * 16), ld_pair_bf16(wr + ks * 16 + 8));
    __nv_bfloat162 lo = __floats2bfloat162_rn(c[0] * 0.125f, c[1] * 0.125f), hi = __floats2bfloat162_rn(c[2] * 0.125f, c[3] * 0.125f);
    *reinterpret_cast<__nv_bfloat162*>(dst + (size_t)gr * 512 + k0 + gc) = lo;
    *reinterpret_cast<__nv_bfloat162*>(dst + (size_t)(gr + 8) * 512 + k0 + gc) = hi;
  }
  if (threadIdx.x < 16) {                                    // score bias: bq . K / 8 (-1e30 masks the padded tokens)
    const int j = threadIdx.x;
    float acc = 0.f;
    if (j < n_mem)
      for (int d = 0; d < 64; ++d) acc = fmaf(__bfloat162float(Kb[(size_t)j * ckv.ld + d]), __ldg(bq[l] + h * 64 + d), acc);
    sbias[(size_t)(l * B + b) * 128 + h * 16 + j] = j < n_mem ? acc * 0.125f : -1e30f;
  }
}

__global__ void __launch_bounds__(128) k_xattn_fold_o(Act ckv, int B, int n_mem, const bf16* __restrict__ woT,
                                                      bf16* __restrict__ Nt) {
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.x, h = blockIdx.y, l = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gr = lane >> 2, gc = (lane & 3) * 2;
  const bf16* Vb = ckv.p + (size_t)(b * n_mem) * ckv.ld + l * 1024 + 512 + h * 64;
  // B fragments: V_b,h as the "col" operand, n = token j, k = d; two n-tiles (j 0..7, 8..15), four k-steps
  uint32_t bv[2][4][2];
#pragma unroll
  for (int n8 = 0; n8 < 2; ++n8) {
    const int j = n8 * 8 + gr;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      bv[n8][ks][0] = j < n_mem ? ld_pair_bf16(Vb + (size_t)j * ckv.ld + ks * 16 + gc) : 0u;
      bv[n8][ks][1] = j < n_mem ? ld_pair_bf16(Vb + (size_t)j * ckv.ld + ks * 16 + gc + 8) : 0u;
    }
  }
  const bf16* W = woT + (size_t)l * 512 * 512 + h * 64;     // [f][h*64 + d]
#pragma unroll 2
  for (int mt = 0; mt < 8; ++mt) {
    const int f0 = warp * 128 + mt * 16;
    float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const bf16* w0 = W + (size_t)(f0 + gr) * 512 + gc;
    const bf16* w1 = w0 + 8 * 512;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4] = {ld_pair_bf16(w0 + ks * 16), ld_pair_bf16(w1 + ks * 16), ld_pair_bf16(w0 + ks * 16 + 8), ld_pair_bf16(w1 + ks * 16 + 8)};
      mma16816(c[0], a, bv[0][ks][0], bv[0][ks][1]);
      mma16816(c[1], a, bv[1][ks][0], bv[1][ks][1]);
    }
#pragma unroll
    for (int n8 = 0; n8 < 2; ++n8) {
      bf16* d0 = Nt + ((size_t)(l * B + b) * 512 + f0 + gr) * 128 + h * 16 + n8 * 8 + gc;
      *reinterpret_cast<__nv_bfloat162*>(d0) = __floats2bfloat162_rn(c[n8][0], c[n8][1]);
      *reinterpret_cast<__nv_bfloat162*>(d0 + 8 * 128) = __floats2bfloat162_rn(c[n8][2], c[n8][3]);
    }
  }
}

__global__ void k_xattn_fence() {}

int launch_xattn_fold(const Act& ckv, int B, int n_mem, int L, const bf16* wq, const float* const* bq, const bf16* woT,
                      bf16* Mt, bf16* Nt, float* sbias, cudaStream_t s) {
  if (n_mem > 16 || ckv.lo) {
    set_last_error("xattn_fold: at most 16 memory tokens, plain bf16 K/V");
    return 1;
  }
  FPNMT_CUDA_OK(launch_k(k_xattn_fold_q, dim3(B, 8, L), dim3(128), 0, s, ckv, B, n_mem, wq, bq, Mt, sbias));
  FPNMT_CUDA_OK(launch_k(k_xattn_fold_o, dim3(B, 8, L), dim3(128), 0, s, ckv, B, n_mem, woT, Nt));
  // A launch WITHOUT the programmatic-serialization attribute: it starts only after the fold kernels have completed
  // and been flushed, so later kernels (which prefetch Mt / Nt before their griddepcontrol.wait) can never overtake them.
  k_xattn_fence<<<1, 32, 0, s>>>();
  FPNMT_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace fpnmt
