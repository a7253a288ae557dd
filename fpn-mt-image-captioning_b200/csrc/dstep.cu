// Cluster-stationary fused decoder kernel (design: dstep.cuh).
//   warp 0      : weight producer - cp.async.bulk of ready-made SWIZZLE_128B tile images through a 6 x 16 KB ring
//   warp 1      : TMEM allocator + tcgen05.mma issuer (M = 128 features x N = 32 rows x K = 16 per instruction)
//   warps 2..17 : workers - operand staging (embedding, LayerNorm), TMEM epilogues, attention, cluster exchanges, beam tail
#include "dstep.cuh"

namespace fpnmt {

constexpr int DS_XB_BYTES = 32768;                      // [32 rows][512] bf16 operand: 8 k-chunks of [32][64] (4 KB, SW128)
constexpr int DS_HB_BYTES = 16384;                      // [32 rows][256] bf16 hidden slice
constexpr int DS_EXTRA_BYTES = 16384;                   // XB | HB | EXTRA = 64 KB: V staging (attention) / selection scratch (tail)
constexpr int DS_OFF_XB = DS_RING * DS_SLOT;
constexpr int DS_OFF_HB = DS_OFF_XB + DS_XB_BYTES;
constexpr int DS_OFF_EXTRA = DS_OFF_HB + DS_HB_BYTES;
constexpr int DS_OFF_Q = DS_OFF_EXTRA + DS_EXTRA_BYTES;   // QS | KS | VS: [32][64] fp32 each
constexpr int DS_OFF_RES = DS_OFF_Q + 3 * 32 * 64 * 4;    // [32][64] fp32 residual slice of this CTA
constexpr int DS_OFF_MISC = DS_OFF_RES + 32 * 64 * 4;
constexpr int DS_MISC_BYTES = 4096;
constexpr int DS_SMEM = DS_OFF_MISC + DS_MISC_BYTES + 1024;
constexpr int DS_W = DS_WORKER_WARPS * 32;                // 512 worker threads
constexpr int DS_CAP = 32;                                // threshold-selection list capacity per (warp, row)
constexpr int DS_NMAX = 16;                               // beam width limit of the fused path
constexpr int NONE_IDX = 0x7fffffff;

size_t dstep_smem_bytes() { return DS_SMEM; }

// ---------------------------------------------------------------------------------------------- small PTX helpers
__device__ __forceinline__ uint32_t ds_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void ds_cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void ds_bulk_g2s_pred(void* smem, const void* g, uint32_t bytes, uint64_t* bar, uint32_t pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
      "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}\n" ::"r"(smem_u32(smem)),
      "l"(g), "r"(bytes), "r"(smem_u32(bar)), "r"(pred)
      : "memory");
}
__device__ __forceinline__ void ds_mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void ds_mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void ds_tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ds_worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(DS_W) : "memory"); }
__device__ __forceinline__ void ds_cp16(uint32_t dst, const void* src, bool ok) {
  const int sz = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ bool ds_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

// ---------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(DS_THREADS, 1) dstep_kernel(const DstepParams p) {
  extern __shared__ uint8_t ds_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ds_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint8_t* XB = smem + DS_OFF_XB;
  uint8_t* HB = smem + DS_OFF_HB;
  uint8_t* SCR = XB;                                               // 64 KB scratch (XB | HB | EXTRA) when no GEMM reads them
  float* QS = reinterpret_cast<float*>(smem + DS_OFF_Q);
  float* KS = QS + 32 * 64;
  float* VS = KS + 32 * 64;
  float* RES = reinterpret_cast<float*>(smem + DS_OFF_RES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DS_OFF_MISC);
  uint64_t* full = bars;                       // [DS_RING]
  uint64_t* empty = bars + DS_RING;            // [DS_RING]
  uint64_t* x_ready = bars + 2 * DS_RING;      // workers -> MMA warp: the B operand of the next GEMM job is in shared memory
  uint64_t* acc_ready = x_ready + 1;           // MMA warp -> workers: the accumulators of the job are complete
  uint64_t* xbar = acc_ready + 1;              // cluster barrier among the workers of the 8 CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 1);
  unsigned* s_rowidx = reinterpret_cast<unsigned*>(smem + DS_OFF_MISC + 256);   // [16 warps][32]
  int* s_small = reinterpret_cast<int*>(smem + DS_OFF_MISC + 256 + 2048);       // 448 ints of small per-phase state

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = (int)ds_cluster_ctarank();
  const int cid = blockIdx.x / DS_CTAS;
  const int nclusters = gridDim.x / DS_CTAS;
  const int N = p.N;
  const int row_base = cid * p.ipc * N;
  const int nrows = min(p.R - row_base, p.ipc * N);        // valid rows of this cluster (whole images)
  const int nimg = nrows / N;

  if (threadIdx.x == 0) {
    for (int s = 0; s < DS_RING; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(x_ready, 1);
    mbar_init(acc_ready, 1);
    mbar_init(xbar, DS_CTAS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ds_cluster_sync_all();            // every CTA's barriers are initialised before any remote arrive

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ weight producer
    const uint32_t leader = elect_one() ? 1u : 0u;
    uint32_t it = 0;
    auto load = [&](const uint8_t* src, uint32_t bytes) {
      const uint32_t slot = it % DS_RING;
      if (it >= DS_RING) mbar_wait(&empty[slot], ((it / DS_RING) & 1) ^ 1);
      mbar_expect_tx_pred(&full[slot], bytes, leader);
      ds_bulk_g2s_pred(ring + slot * DS_SLOT, src, bytes, &full[slot], leader);
      ++it;
    };
    for (int s = 0; s < p.nsteps; ++s) {
      for (int l = 0; l < p.L; ++l) {
        const uint8_t* w = p.wstream + ((size_t)l * DS_CTAS + cta) * DS_LAYER_STREAM;
        for (int i = 0; i < 8; ++i, w += 16384) load(w, 16384);          // [q_c | k_c]
        for (int i = 0; i < 32; ++i, w += 8192) load(w, 8192);           // v_c, o1, q2, o2
        for (int i = 0; i < 32; ++i, w += 16384) load(w, 16384);         // ffn1 (2 tiles x 8), ffn2 (4 tiles x 4)
      }
      const uint8_t* w = p.wstream + p.final_off + (size_t)cta * p.ntv * 8 * 16384;
      for (int i = 0; i < p.ntv * 8; ++i, w += 16384) load(w, 16384);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    const uint32_t leader = elect_one() ? 1u : 0u;
    constexpr uint32_t IDESC = umma_idesc_bf16(128, DS_ROWS);
    const uint32_t ring_a = smem_u32(ring), xb_a = smem_u32(XB), hb_a = smem_u32(HB);
    uint32_t it = 0, job = 0;
    auto tile = [&](uint32_t acc_col, int nchunks, uint32_t b_addr) {
      for (int kc = 0; kc < nchunks; ++kc) {
        const uint32_t slot = it % DS_RING;
        mbar_wait(&full[slot], (it / DS_RING) & 1);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(ring_a + slot * DS_SLOT);
        const uint64_t bdesc = umma_desc_sw128(b_addr + kc * 4096);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_pred(tmem_base + acc_col, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC, (kc > 0 || k > 0) ? 1u : 0u, leader);
        umma_commit_pred(&empty[slot], leader);
        ++it;
      }
    };
    auto begin_job = [&]() {
      mbar_wait(x_ready, job & 1);
      tc_fence_after();
    };
    auto end_job = [&]() {
      umma_commit_pred(acc_ready, leader);
      ++job;
    };
    for (int s = 0; s < p.nsteps; ++s) {
      for (int l = 0; l < p.L; ++l) {
        begin_job(); tile(0, 8, xb_a); tile(32, 8, xb_a); end_job();                       // qkv
        begin_job(); tile(0, 8, xb_a); end_job();                                          // o1
        begin_job(); tile(0, 8, xb_a); end_job();                                          // q2
        begin_job(); tile(0, 8, xb_a); end_job();                                          // o2
        begin_job(); tile(0, 8, xb_a); tile(32, 8, xb_a); end_job();                       // ffn1
        begin_job();
        for (int ft = 0; ft < 4; ++ft) tile(64 + 32 * ft, 4, hb_a);                        // ffn2 split-K partials
        end_job();
      }
      begin_job();
      for (int vt = 0; vt < p.ntv; ++vt) tile(32 * vt, 8, xb_a);                           // vocabulary slice
      end_job();
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------------ workers
    const int ww = warp - 2;                       // worker warp 0..15
    const int wt = ww * 32 + lane;                 // worker thread 0..511
    const int q = warp & 3;                        // TMEM lane quarter of this warp
    const int g = ww >> 2;                         // column (row) group: rows g*8 .. g*8+7
    const int r16 = wt >> 4, p16 = wt & 15;        // (row, 1/16 of the row) mapping
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 8);
    uint32_t job = 0, xphase = 0;
    const size_t xrow0 = (size_t)cid * 32;         // first row of this cluster in the exchange buffers

    auto signal_x_ready = [&]() {                  // every worker has written its part of the operand
      fence_proxy_async();
      tc_fence_before();
      ds_worker_sync();
      if (wt == 0) mbar_arrive(x_ready);
    };
    auto wait_acc = [&]() {
      mbar_wait(acc_ready, job & 1);
      tc_fence_after();
      ++job;
    };
    auto xsync = [&]() {                           // cluster barrier among the workers of the 8 CTAs (release / acquire)
      ds_worker_sync();
      if (wt < DS_CTAS) {
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
        ds_mbar_arrive_remote(xbar, (uint32_t)wt);
      }
      ds_mbar_wait_cluster(xbar, xphase);
      xphase ^= 1;
    };
    // x[32] = features p16*32 .. +31 of row r16 -> bf16 operand (SW128 K-major) + this CTA's fp32 residual slice
    auto store_operand = [&](const float* x) {
      const int kc = p16 >> 1, j0 = (p16 & 1) * 4;
      uint8_t* base = XB + kc * 4096 + r16 * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(base + (((j0 + i) ^ (r16 & 7)) << 4)) = pack8(x + 8 * i);
      if (kc == cta) {
        float* rs = RES + r16 * 64 + (p16 & 1) * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(rs + 4 * i) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
      }
    };
    // LayerNorm (eps 1e-6) of the exchanged pre-activation rows; every CTA normalises all 32 rows itself
    auto ln_load = [&](const float* gam, const float* bet, int l, int which) {
      float x[32];
      const float* src = p.x_pre + (xrow0 + r16) * 512 + p16 * 32;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src) + i);
        x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) s += x[i];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * (1.f / 512.f);
      float m2 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float d = x[i] - mean;
        m2 = fmaf(d, d, m2);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
      const float rstd = rsqrtf(m2 * (1.f / 512.f) + 1e-6f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gam + p16 * 32) + i);
        const float4 be = __ldg(reinterpret_cast<const float4*>(bet + p16 * 32) + i);
        x[4 * i] = (x[4 * i] - mean) * rstd * ga.x + be.x;
        x[4 * i + 1] = (x[4 * i + 1] - mean) * rstd * ga.y + be.y;
        x[4 * i + 2] = (x[4 * i + 2] - mean) * rstd * ga.z + be.z;
        x[4 * i + 3] = (x[4 * i + 3] - mean) * rstd * ga.w + be.w;
      }
      store_operand(x);
      if (p.dbg && cta == 0) {
        float* d = p.dbg + (((size_t)(l * 3 + which) * nclusters * 32) + xrow0 + r16) * 512 + p16 * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) d[i] = x[i];
      }
    };
    auto load_att_operand = [&]() {                // x_att rows (bf16) -> operand, no conversion
      const uint4* src = reinterpret_cast<const uint4*>(p.x_att + (xrow0 + r16) * 512 + p16 * 32);
      const int kc = p16 >> 1, j0 = (p16 & 1) * 4;
      uint8_t* base = XB + kc * 4096 + r16 * 128;
      uint4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = __ldcg(src + i);
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(base + (((j0 + i) ^ (r16 & 7)) << 4)) = v[i];
    };
    // 64-feature Dense slice epilogue: acc + bias + residual slice -> exchanged pre-LayerNorm rows
    auto epi_slice = [&](const float* bias) {
      if (q < 2) {
        float v[8];
        ds_tmem_ld8(t_lane, v);
        const int f = q * 32 + lane;
        const float b = __ldg(bias + cta * 64 + f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int row = g * 8 + j;
          p.x_pre[(xrow0 + row) * 512 + cta * 64 + f] = v[j] + b + RES[row * 64 + f];
        }
      }
    };

    for (int s = 0; s < p.nsteps; ++s) {
      const int t = p.t0 + s;
      const int cur = t & 1, nxt = cur ^ 1;
      // ---- decoder input: x = embedding[token] + pos[t]  (transformer.py:326-329; no sqrt(d) scaling, :327 is commented out)
      {
        float x[32];
        if (r16 < nrows) {
          const int tok = __ldcg(p.st.last_tok + row_base + r16);
          const float4* e = reinterpret_cast<const float4*>(p.emb + (size_t)tok * 512 + p16 * 32);
          const float4* ps = reinterpret_cast<const float4*>(p.pos + (size_t)t * 512 + p16 * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a = __ldg(e + i), b = __ldg(ps + i);
            x[4 * i] = a.x + b.x; x[4 * i + 1] = a.y + b.y; x[4 * i + 2] = a.z + b.z; x[4 * i + 3] = a.w + b.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = 0.f;
        }
        store_operand(x);
      }
      signal_x_ready();

      for (int l = 0; l < p.L; ++l) {
        const float* lp = p.lparams + (size_t)l * DSB_SIZE;
        // ================================================================== qkv epilogue -> self-attention of head `cta`
        wait_acc();
        {
          float v[8];
          ds_tmem_ld8(t_lane, v);                                   // tile [q_c | k_c]
          const int f = q * 32 + lane;
          const float b = __ldg(lp + (q < 2 ? DSB_Q + cta * 64 + f : DSB_K + cta * 64 + f - 64));
          float* dst = (q < 2 ? QS : KS) + (f & 63);
          const float sc = q < 2 ? 0.125f : 1.f;                    // 1 / sqrt(depth) folded into q (transformer.py:91-92)
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[(g * 8 + j) * 64] = (v[j] + b) * sc;
          if (q < 2) {
            ds_tmem_ld8(t_lane + 32, v);                            // tile [v_c]
            const float bv = __ldg(lp + DSB_V + cta * 64 + f);
#pragma unroll
            for (int j = 0; j < 8; ++j) VS[(g * 8 + j) * 64 + f] = v[j] + bv;
          }
        }
        tc_fence_before();
        ds_worker_sync();
        // append K/V of position t to the cache (bf16)
        if (r16 < nrows) {
          const float* src = (p16 < 8 ? KS : VS) + r16 * 64 + (p16 & 7) * 8;
          bf16* dstc = (p16 < 8 ? p.kcache : p.vcache) + (((size_t)l * p.R + row_base + r16) * p.T + t) * 512 + cta * 64 + (p16 & 7) * 8;
          *reinterpret_cast<uint4*>(dstc) = pack8(src);
        }
        // self-attention (transformer.py:88-102 with the causal mask implicit): warp -> rows 2*ww, 2*ww+1
        {
          uint4* vst = reinterpret_cast<uint4*>(SCR + ww * 4096);       // V rows of the current 32-position chunk
          unsigned* rowi = s_rowidx + ww * 32;
          const int pg = lane >> 3, dg = lane & 7;
          const bf16* kc_l = p.kcache + (size_t)l * p.R * p.T * 512 + cta * 64;
          const bf16* vc_l = p.vcache + (size_t)l * p.R * p.T * 512 + cta * 64;
          for (int rr = 0; rr < 2; ++rr) {
            const int lr = ww * 2 + rr;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = 0.f;
            float m = -INFINITY, lsum = 0.f;
            if (lr < nrows) {
              const int* anc = p.st.anc[cur] + (size_t)(row_base + lr) * p.T;
              float q8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) q8[j] = QS[lr * 64 + dg * 8 + j];
              for (int k0 = 0; k0 < t; k0 += 32) {
                const int pos = k0 + lane;
                const int kmax = min(32, t - k0);
                rowi[lane] = (pos < t) ? (unsigned)(__ldcg(anc + pos) * p.T + pos) : 0u;
                __syncwarp();
                {
                  const uint32_t d0 = smem_u32(vst);
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const int u = i * 32 + lane, r = u >> 3;
                    const bool ok = r < kmax;
                    ds_cp16(d0 + u * 16, vc_l + (size_t)rowi[ok ? r : 0] * 512 + (u & 7) * 8, ok);
                  }
                  asm volatile("cp.async.commit_group;" ::: "memory");
                }
                float sc[8];
                {
                  uint4 kh[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const int pp = pg + 4 * i;
                    kh[i] = __ldcg(reinterpret_cast<const uint4*>(kc_l + (size_t)rowi[pp < kmax ? pp : 0] * 512 + dg * 8));
                  }
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    float f[8];
                    unpack8(kh[i], f);
                    float a = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) a = fmaf(q8[j], f[j], a);
                    sc[i] = a;
                  }
                }
                float cmax = -INFINITY;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  sc[i] += __shfl_xor_sync(0xffffffffu, sc[i], 1);
                  sc[i] += __shfl_xor_sync(0xffffffffu, sc[i], 2);
                  sc[i] += __shfl_xor_sync(0xffffffffu, sc[i], 4);
                  if (pg + 4 * i >= kmax) sc[i] = -INFINITY;
                  cmax = fmaxf(cmax, sc[i]);
                }
                cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, 8));
                cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, 16));
                const float mn = fmaxf(m, cmax);
                const float corr = __expf(m - mn);
                float psum = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  sc[i] = (pg + 4 * i < kmax) ? __expf(sc[i] - mn) : 0.f;
                  psum += sc[i];
                }
                psum += __shfl_xor_sync(0xffffffffu, psum, 8);
                psum += __shfl_xor_sync(0xffffffffu, psum, 16);
                lsum = lsum * corr + psum;
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] *= corr;
                m = mn;
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int pp = pg + 4 * i;
                  if (pp < kmax) {
                    float f[8];
                    unpack8(vst[pp * 8 + dg], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = fmaf(sc[i], f[j], o[j]);
                  }
                }
                __syncwarp();
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {          // sum the four position groups: lanes 0..7 hold dims lane*8 .. +7
                o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
                o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
              }
              // the new position itself (K/V in shared memory, fp32)
              const float scn = warp_sum(QS[lr * 64 + 2 * lane] * KS[lr * 64 + 2 * lane] + QS[lr * 64 + 2 * lane + 1] * KS[lr * 64 + 2 * lane + 1]);
              const float mn = fmaxf(m, scn);
              const float corr = __expf(m - mn);
              const float pn = __expf(scn - mn);
              lsum = lsum * corr + pn;
              const float inv = 1.f / lsum;
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = (o[i] * corr + pn * VS[lr * 64 + (lane & 7) * 8 + i]) * inv;
            }
            if (lane < 8) *reinterpret_cast<uint4*>(p.x_att + (xrow0 + lr) * 512 + cta * 64 + lane * 8) = pack8(o);
          }
        }
        xsync();                                                    // A: attention heads of all CTAs
        load_att_operand();
        signal_x_ready();
        // ================================================================== o1 + residual -> LayerNorm1
        wait_acc();
        epi_slice(lp + DSB_O1);
        tc_fence_before();
        xsync();                                                    // B
        ln_load(lp + DSB_LN1G, lp + DSB_LN1B, l, 0);
        signal_x_ready();
        // ================================================================== q2 -> cross-attention over the memory tokens
        wait_acc();
        if (q < 2) {
          float v[8];
          ds_tmem_ld8(t_lane, v);
          const int f = q * 32 + lane;
          const float b = __ldg(lp + DSB_Q2 + cta * 64 + f);
#pragma unroll
          for (int j = 0; j < 8; ++j) QS[(g * 8 + j) * 64 + f] = (v[j] + b) * 0.125f;
        }
        tc_fence_before();
        ds_worker_sync();
        {
          const int img = min((row_base + r16) / N, p.B - 1);
          const int j = p16;                                        // memory token of this thread
          const bf16* kv = p.ckv + (size_t)img * p.n_mem * p.ckv_ld + (size_t)l * 1024 + cta * 64;
          float sc = -INFINITY;
          if (j < p.n_mem) {
            const uint4* kr = reinterpret_cast<const uint4*>(kv + (size_t)j * p.ckv_ld);
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float f[8];
              unpack8(__ldg(kr + i), f);
#pragma unroll
              for (int d = 0; d < 8; ++d) a = fmaf(QS[r16 * 64 + i * 8 + d], f[d], a);
            }
            sc = a;
          }
          float mx = sc;
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          float pr = (j < p.n_mem) ? __expf(sc - mx) : 0.f;
          float sm = pr;
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
          pr /= sm;
          float o4[4] = {0.f, 0.f, 0.f, 0.f};
          const bf16* vr = kv + 512 + p16 * 4;                      // this thread's 4 output dims
          for (int jj = 0; jj < p.n_mem; ++jj) {
            const float pj = __shfl_sync(0xffffffffu, pr, (lane & 16) | jj);
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(vr + (size_t)jj * p.ckv_ld));
            o4[0] = fmaf(pj, __uint_as_float(u.x << 16), o4[0]);
            o4[1] = fmaf(pj, __uint_as_float(u.x & 0xffff0000u), o4[1]);
            o4[2] = fmaf(pj, __uint_as_float(u.y << 16), o4[2]);
            o4[3] = fmaf(pj, __uint_as_float(u.y & 0xffff0000u), o4[3]);
          }
          *reinterpret_cast<uint2*>(p.x_att + (xrow0 + r16) * 512 + cta * 64 + p16 * 4) = make_uint2(pack2(o4[0], o4[1]), pack2(o4[2], o4[3]));
        }
        xsync();                                                    // C
        load_att_operand();
        signal_x_ready();
        // ================================================================== o2 + residual -> LayerNorm2
        wait_acc();
        epi_slice(lp + DSB_O2);
        tc_fence_before();
        xsync();                                                    // D
        ln_load(lp + DSB_LN2G, lp + DSB_LN2B, l, 1);
        signal_x_ready();
        // ================================================================== ffn1 + LeakyReLU(0.2) -> hidden slice operand
        wait_acc();
#pragma unroll
        for (int ft = 0; ft < 2; ++ft) {
          float v[8];
          ds_tmem_ld8(t_lane + 32 * ft, v);
          const int hf = ft * 128 + q * 32 + lane;
          const float b = __ldg(lp + DSB_F1 + cta * 256 + hf);
          uint8_t* base = HB + (hf >> 6) * 4096 + (hf & 7) * 2;
          const int c16 = (hf & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = g * 8 + j;
            float y = v[j] + b;
            y = y >= 0.f ? y : 0.2f * y;
            *reinterpret_cast<bf16*>(base + row * 128 + ((c16 ^ (row & 7)) << 4)) = __float2bfloat16_rn(y);
          }
        }
        signal_x_ready();
        // ================================================================== ffn2 split-K partials -> reduce -> LayerNorm3
        wait_acc();
#pragma unroll
        for (int ft = 0; ft < 4; ++ft) {
          float v[8];
          ds_tmem_ld8(t_lane + 64 + 32 * ft, v);
          float* dst = p.x_part + (((size_t)cid * DS_CTAS + cta) * 32 + g * 8) * 512 + ft * 128 + q * 32 + lane;
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j * 512] = v[j];
        }
        tc_fence_before();
        xsync();                                                    // E
        {
          float4 a = __ldg(reinterpret_cast<const float4*>(lp + DSB_F2 + cta * 64 + p16 * 4));
          const float4 rs = *reinterpret_cast<const float4*>(RES + r16 * 64 + p16 * 4);
          a.x += rs.x; a.y += rs.y; a.z += rs.z; a.w += rs.w;
#pragma unroll
          for (int c2 = 0; c2 < DS_CTAS; ++c2) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(p.x_part + (((size_t)cid * DS_CTAS + c2) * 32 + r16) * 512 + cta * 64 + p16 * 4));
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
          }
          *reinterpret_cast<float4*>(p.x_pre + (xrow0 + r16) * 512 + cta * 64 + p16 * 4) = a;
        }
        xsync();                                                    // F
        ln_load(lp + DSB_LN3G, lp + DSB_LN3B, l, 2);
        signal_x_ready();
      }

      // ==================================================================== vocabulary projection epilogue + beam tail
      wait_acc();
      const int vbase = cta * p.vslice + q * 32 + lane;             // vocabulary id of this thread in tile 0
      if (p.mode == 1) {
        for (int vt = 0; vt < p.ntv; ++vt) {
          float v[8];
          ds_tmem_ld8(t_lane + 32 * vt, v);
          const int vid = vbase + vt * 128;
          const float b = __ldg(p.vbias + vid);
          if (vid < p.V) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (g * 8 + j < nrows) p.logits_out[(size_t)(row_base + g * 8 + j) * p.ld_logits + vid] = v[j] + b;
          }
        }
        tc_fence_before();
        ds_worker_sync();
        continue;
      }
      // ---- per (warp, row): maxima, selection threshold, sum of exponentials, top-N of the warp's 32 x ntv logits
      float* s_lv = reinterpret_cast<float*>(SCR);                   // [16][8][CAP]
      int* s_li = reinterpret_cast<int*>(SCR + 16384);               // [16][8][CAP]
      float* s_wv = reinterpret_cast<float*>(SCR + 32768);           // [16][8][DS_NMAX] warp-level winners
      int* s_wi = reinterpret_cast<int*>(SCR + 32768 + 8192);        // [16][8][DS_NMAX]
      float* s_wm = reinterpret_cast<float*>(SCR + 49152);           // [16][8] warp max
      float* s_ws = s_wm + 128;                                      // [16][8] warp sum-exp
      int* s_cnt = s_small;                                          // [16][8]
      {
        float lmax[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) lmax[j] = -INFINITY;
        for (int vt = 0; vt < p.ntv; ++vt) {
          float v[8];
          ds_tmem_ld8(t_lane + 32 * vt, v);
          const int vid = vbase + vt * 128;
          if (vid < p.V) {
            const float b = __ldg(p.vbias + vid);
#pragma unroll
            for (int j = 0; j < 8; ++j) lmax[j] = fmaxf(lmax[j], v[j] + b);
          }
        }
        float tau[8], wmax[8];
        const int nsel = min(N, 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gmx = lmax[j];
          int rank = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float gi = __shfl_sync(0xffffffffu, gmx, i);
            rank += (gi > gmx || (gi == gmx && i < lane)) ? 1 : 0;
          }
          const unsigned hit = __ballot_sync(0xffffffffu, rank == nsel - 1);
          tau[j] = __shfl_sync(0xffffffffu, gmx, __ffs(hit) - 1);
          wmax[j] = warp_max(gmx);
        }
        if (lane < 8) s_cnt[ww * 8 + lane] = 0;
        __syncwarp();
        float sum[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[j] = 0.f;
        for (int vt = 0; vt < p.ntv; ++vt) {
          float v[8];
          ds_tmem_ld8(t_lane + 32 * vt, v);
          const int vid = vbase + vt * 128;
          if (vid < p.V) {
            const float b = __ldg(p.vbias + vid);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = v[j] + b;
              sum[j] += __expf(x - wmax[j]);
              if (x >= tau[j]) {
                const int pos = atomicAdd(&s_cnt[ww * 8 + j], 1);
                if (pos < DS_CAP) {
                  s_lv[(ww * 8 + j) * DS_CAP + pos] = x;
                  s_li[(ww * 8 + j) * DS_CAP + pos] = vid;
                }
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float sj = warp_sum(sum[j]);
          if (lane == 0) {
            s_wm[ww * 8 + j] = wmax[j];
            s_ws[ww * 8 + j] = (wmax[j] > -INFINITY) ? sj : 0.f;
          }
        }
        __syncwarp();
        for (int j = 0; j < 8; ++j) {
          const int cnt = s_cnt[ww * 8 + j];
          float* wv = s_wv + (ww * 8 + j) * DS_NMAX;
          int* wi = s_wi + (ww * 8 + j) * DS_NMAX;
          if (cnt <= DS_CAP) {                       // rank the few survivors by counting (value desc, id asc)
            const float* lv = s_lv + (ww * 8 + j) * DS_CAP;
            const int* li = s_li + (ww * 8 + j) * DS_CAP;
            if (lane < cnt) {
              const float v = lv[lane];
              const int id = li[lane];
              int rank = 0;
              for (int e = 0; e < cnt; ++e) rank += ds_better(lv[e], li[e], v, id) ? 1 : 0;
              if (rank < N) {
                wv[rank] = v;
                wi[rank] = id;
              }
            }
            if (lane >= cnt && lane < N) {
              wv[lane] = -INFINITY;
              wi[lane] = NONE_IDX;
            }
          } else {
            // exact fallback for massive ties: N rounds of "best element strictly after the previous winner"
            float pv = INFINITY;
            int pi = -1;
            for (int k = 0; k < N; ++k) {
              float bv = -INFINITY;
              int bi = NONE_IDX;
              for (int vt = 0; vt < p.ntv; ++vt) {
                float v[8];
                ds_tmem_ld8(t_lane + 32 * vt, v);
                const int vid = vbase + vt * 128;
                if (vid < p.V) {
                  float x = 0.f;
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj)
                    if (jj == j) x = v[jj];
                  x += __ldg(p.vbias + vid);
                  const bool after = (x < pv) || (x == pv && vid > pi);
                  if (after && ds_better(x, vid, bv, bi)) {
                    bv = x;
                    bi = vid;
                  }
                }
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ds_better(ov, oi, bv, bi)) {
                  bv = ov;
                  bi = oi;
                }
              }
              if (lane == 0) {
                wv[k] = bv;
                wi[k] = bi;
              }
              pv = bv;
              pi = bi;
            }
          }
        }
      }
      tc_fence_before();
      ds_worker_sync();
      // ---- CTA level: merge the 4 quarter-warps of every row -> exchanged per-CTA candidates and (max, sum-exp)
      {
        const int gr = r16 >> 3, jr = r16 & 7;      // the 4 warps gr*4 .. gr*4+3 hold row r16 (as their local row jr)
        const int n4 = 4 * N;
        for (int e = p16; e < n4; e += 16) {
          const int wq = gr * 4 + e / N, k = e % N;
          const float v = s_wv[(wq * 8 + jr) * DS_NMAX + k];
          const int id = s_wi[(wq * 8 + jr) * DS_NMAX + k];
          int rank = 0;
          for (int e2 = 0; e2 < n4; ++e2) {
            const int w2 = gr * 4 + e2 / N, k2 = e2 % N;
            const float v2 = s_wv[(w2 * 8 + jr) * DS_NMAX + k2];
            const int i2 = s_wi[(w2 * 8 + jr) * DS_NMAX + k2];
            rank += (v2 > v || (v2 == v && (i2 < id || (i2 == id && e2 < e)))) ? 1 : 0;
          }
          if (rank < N) {
            p.x_cval[((xrow0 + r16) * DS_CTAS + cta) * N + rank] = v;
            p.x_cidx[((xrow0 + r16) * DS_CTAS + cta) * N + rank] = id;
          }
        }
        if (p16 == 0) {
          float M = -INFINITY;
#pragma unroll
          for (int i = 0; i < 4; ++i) M = fmaxf(M, s_wm[(gr * 4 + i) * 8 + jr]);
          float S = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float mi = s_wm[(gr * 4 + i) * 8 + jr];
            if (mi > -INFINITY) S += s_ws[(gr * 4 + i) * 8 + jr] * __expf(mi - M);
          }
          p.x_stat[((xrow0 + r16) * DS_CTAS + cta) * 2] = M;
          p.x_stat[((xrow0 + r16) * DS_CTAS + cta) * 2 + 1] = S;
        }
      }
      xsync();                                                      // T1: candidates of all vocabulary slices
      // ---- per image (CTA i handles images i, i+8, ..): row top-N over the 8 slices, then N x N -> N beams
      for (int im = cta; im < nimg; im += DS_CTAS) {
        float* m_v = reinterpret_cast<float*>(SCR);                  // [N][8N] logits
        int* m_i = reinterpret_cast<int*>(SCR + 8192);               // [N][8N] vocabulary ids
        float* r_c = reinterpret_cast<float*>(SCR + 16384);          // [N][N] row winners: candidate scores
        int* r_i = reinterpret_cast<int*>(SCR + 16384 + 1024);       // [N][N] ids
        float* s_lse = reinterpret_cast<float*>(SCR + 16384 + 2048); // [N]
        int* s_parent = s_small + 128;                               // [N]
        int* s_token = s_small + 160;                                // [N]
        const int n8 = 8 * N;
        const int rows0 = row_base + im * N;                         // first global row of the image
        const size_t xr0 = xrow0 + (size_t)im * N;
        ds_worker_sync();                                            // scratch reuse across the image loop
        for (int e = wt; e < N * n8; e += DS_W) {
          const int n = e / n8, k = e % n8;
          m_v[e] = __ldcg(p.x_cval + (xr0 + n) * n8 + k);
          m_i[e] = __ldcg(p.x_cidx + (xr0 + n) * n8 + k);
        }
        if (wt < N) {
          float M = -INFINITY;
          float ms[DS_CTAS], ss[DS_CTAS];
#pragma unroll
          for (int c2 = 0; c2 < DS_CTAS; ++c2) {
            ms[c2] = __ldcg(p.x_stat + ((xr0 + wt) * DS_CTAS + c2) * 2);
            ss[c2] = __ldcg(p.x_stat + ((xr0 + wt) * DS_CTAS + c2) * 2 + 1);
            M = fmaxf(M, ms[c2]);
          }
          float S = 0.f;
#pragma unroll
          for (int c2 = 0; c2 < DS_CTAS; ++c2)
            if (ms[c2] > -INFINITY) S += ss[c2] * __expf(ms[c2] - M);
          s_lse[wt] = M + logf(S);
        }
        ds_worker_sync();
        for (int e = wt; e < N * n8; e += DS_W) {
          const int n = e / n8, k = e % n8;
          const float v = m_v[e];
          const int id = m_i[e];
          int rank = 0;
          for (int k2 = 0; k2 < n8; ++k2) {
            const float v2 = m_v[n * n8 + k2];
            const int i2 = m_i[n * n8 + k2];
            rank += (v2 > v || (v2 == v && (i2 < id || (i2 == id && k2 < k)))) ? 1 : 0;
          }
          if (rank < N) {
            // candidate score of pipeline.py:117-123 in the log domain: beam log-prob + log_softmax(logit)
            const float score = __ldcg(p.st.score[cur] + rows0 + n);
            r_c[n * N + rank] = (id != NONE_IDX) ? score + (v - s_lse[n]) : -INFINITY;
            r_i[n * N + rank] = (id != NONE_IDX) ? n * p.V + id : NONE_IDX;        // flat index over the N x V candidates
          }
        }
        ds_worker_sync();
        for (int e = wt; e < N * N; e += DS_W) {
          const float c = r_c[e];
          const int f = r_i[e];
          if (f == NONE_IDX) continue;
          int rank = 0;
          for (int e2 = 0; e2 < N * N; ++e2) rank += ds_better(r_c[e2], r_i[e2], c, f) ? 1 : 0;
          if (rank < N) {                                            // new beam `rank` (pipeline.py:127-131, 140-141)
            const int par = f / p.V, tok = f - par * p.V;
            s_parent[rank] = par;
            s_token[rank] = tok;
            p.st.score[nxt][rows0 + rank] = c;
            p.st.last_tok[rows0 + rank] = tok;
            if (p.st.parent_out) p.st.parent_out[(size_t)t * p.st.Btot * N + rows0 + rank] = par;
            if (p.st.token_out) p.st.token_out[(size_t)t * p.st.Btot * N + rows0 + rank] = tok;
            if (rank == 0 && p.st.step_logprob) p.st.step_logprob[(size_t)t * p.st.Btot + rows0 / N] = c;
          }
        }
        ds_worker_sync();
        const int T = p.T;
        for (int n = ww; n < N; n += DS_WORKER_WARPS) {              // reorder sequences + ancestry (pipeline.py:134-137)
          const int par = s_parent[n], tok = s_token[n];
          const int* sseq = p.st.seq[cur] + (size_t)(rows0 + par) * (T + 1);
          int* dseq = p.st.seq[nxt] + (size_t)(rows0 + n) * (T + 1);
          for (int j = lane; j <= t; j += 32) dseq[j] = __ldcg(sseq + j);
          const int* sanc = p.st.anc[cur] + (size_t)(rows0 + par) * T;
          int* danc = p.st.anc[nxt] + (size_t)(rows0 + n) * T;
          for (int j = lane; j < t; j += 32) danc[j] = __ldcg(sanc + j);
          if (lane == 0) {
            dseq[t + 1] = tok;
            danc[t] = rows0 + par;
          }
        }
        ds_worker_sync();
        if (ww == 0) {                                               // top beam is rank 0 (pipeline.py:143-148)
          const int b = rows0 / N;
          const int top_tok = s_token[0];
          if (!p.st.done[b] && (top_tok == p.st.end_id || t == T - 1)) {
            const int* res = p.st.seq[nxt] + (size_t)rows0 * (T + 1);
            const int len = (top_tok == p.st.end_id) ? t : t + 1;    // strip <start> and a trailing <end>
            for (int j = lane; j < len; j += 32) p.st.out_ids[(size_t)b * T + j] = res[1 + j];
            __syncwarp();
            if (lane == 0) {
              p.st.out_len[b] = len;
              p.st.done[b] = 1;
              atomicAdd(p.st.n_done, 1);
            }
          }
        }
      }
      xsync();                                                      // T2: next tokens / ancestry / scores of the cluster
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  ds_cluster_sync_all();
}

int dstep_set_attributes() {
  FPNMT_CUDA_OK(cudaFuncSetAttribute(dstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DS_SMEM));
  return 0;
}

int dstep_launch(const DstepParams& p, cudaStream_t stream) {
  if (p.N < 1 || p.N > DS_NMAX || p.ntv < 1 || p.ntv > DS_MAX_VTILES || p.n_mem < 1 || p.n_mem > 16 || p.nsteps < 1 ||
      p.t0 < 0 || p.t0 + p.nsteps > p.T) {
    set_last_error("dstep_launch: configuration outside the fused decoder's limits");
    return 1;
  }
  const int clusters = (p.B + p.ipc - 1) / p.ipc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * DS_CTAS);
  cfg.blockDim = dim3(DS_THREADS);
  cfg.dynamicSmemBytes = DS_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = DS_CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FPNMT_CUDA_OK(cudaLaunchKernelEx(&cfg, dstep_kernel, p));
  return 0;
}

}  // namespace fpnmt
