// tcgen05 implicit-GEMM kernel (see igemm.cuh).  Warp-specialised, persistent, 320 threads:
//   warp 0      : TMA producer (one lane): global -> smem ring (A box + B box per stage).  The weight (B) boxes of the
//                 first ring pass are issued BEFORE griddepcontrol.wait (weights are static), so under programmatic
//                 dependent launch they stream in while the previous kernel is still running.
//   warp 1      : TMEM allocator + MMA issuer (one lane): tcgen05.mma 128 x BN x 16, fp32 accumulators in TMEM.
//   warps 2..9  : epilogue.  Warp e owns TMEM lane quarter (warp_id % 4) and one half of the tile's columns.
//                 Per 32-column unit: tcgen05.ld -> registers; + bias; + residual (prefetched with cp.async into a
//                 swizzled per-warp staging buffer while the MMAs of the tile are still running); activation; bf16
//                 pack -> staging buffer -> TRANSPOSED 16-byte global stores (8 rows x 64 contiguous bytes per warp
//                 instruction: full 32 B sectors, instead of 32 rows x 16 B).
// Two TMEM accumulator buffers (2*BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#include "igemm.cuh"

namespace fpnmt {

constexpr int A_STAGE_BYTES = IG_BM * IG_BK * 2;   // 16 KB
constexpr int EPI_WARPS = 8;
constexpr int UNIT_BYTES = 32 * 64;                // 32 rows x 32 bf16 columns
constexpr int IG_SMEM_MAX = 226 * 1024;            // opt-in limit per CTA on sm_100 is 227 KB; 1 KB left for static shared memory (debug builds)

__host__ __device__ constexpr int ig_stages(int BN) { return BN == 256 ? 3 : (BN == 128 ? 5 : (BN == 64 ? 6 : 8)); }
__host__ __device__ constexpr int ig_b_bytes(int BN) { return BN * IG_BK * 2; }
__host__ __device__ constexpr int ig_units_per_warp(int BN) { return BN >= 64 ? BN / 64 : 1; }

int igemm_stages(int BN) { return ig_stages(BN); }
size_t igemm_smem_bytes(int BN, int b_chunks) {
  const int nb = b_chunks > 0 ? b_chunks : ig_stages(BN);
  return (size_t)ig_stages(BN) * A_STAGE_BYTES + (size_t)nb * ig_b_bytes(BN) + (size_t)EPI_WARPS * ig_units_per_warp(BN) * UNIT_BYTES +
         (size_t)EPI_WARPS * 32 * 2 * sizeof(long long) + 1024 /*align slack*/ + 256 /*barriers*/;
}

__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int n) {
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
  }
}
// staging buffer addressing: row r (0..31) holds 4 chunks of 16 B; chunk c is stored at slot c ^ ((r >> 1) & 3) so that
// both "thread = row" and "8 rows x 4 chunks" access patterns are bank-conflict free.
__device__ __forceinline__ uint4* stage_ptr(uint8_t* buf, int r, int c) {
  return reinterpret_cast<uint4*>(buf + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
}

template <int BN>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
             const __grid_constant__ CUtensorMap tmB, const IgemmParams p) {
  constexpr int STAGES = ig_stages(BN);
  constexpr int B_STAGE_BYTES = ig_b_bytes(BN);
  constexpr int TMEM_COLS = 2 * BN;
  constexpr int UPW = ig_units_per_warp(BN);       // 32-column units per epilogue warp
  constexpr int UNITS = BN / 32;
  constexpr uint32_t IDESC = umma_idesc_bf16(IG_BM, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  const int taps = p.taps_y * p.taps_x;
  const int kiters = p.nterms * taps * p.kchunks;
  // Stationary weights (p.b_stat): with one output-channel tile every tile of this CTA multiplies by the same [BN][K] panel; it is
  // fetched ONCE (before the grid dependency resolves) instead of once per tile, and the ring carries only activation chunks.
  // res2*_2b (3x3, 64 -> 64 channels): 72 KB of the 216 KB a tile used to pull through TMA - the layer is ingest-bound.
  const int nB = p.b_stat ? kiters : STAGES;
  uint8_t* sStage = smem + STAGES * A_STAGE_BYTES + nB * B_STAGE_BYTES;               // [EPI_WARPS][UPW][UNIT_BYTES]
  long long* sOff = reinterpret_cast<long long*>(sStage + EPI_WARPS * UPW * UNIT_BYTES);   // [EPI_WARPS][2][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOff + EPI_WARPS * 64);
  uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* bstat_bar = bars + 2 * STAGES + 5;  // stationary weight panel resident

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

#ifdef FPNMT_DBG_STAMPS   // build.py --dbg-stamps: globaltimer timeline of block 0 (FPNMT_DBG_OP=<op name>); off in product builds
  __shared__ long long* s_dbg;
  if (threadIdx.x == 0) {
    s_dbg = nullptr;
    if (p.dbg && blockIdx.x == 0) {
      const long long inst = (long long)atomicAdd((unsigned long long*)p.dbg, 1ull);
      s_dbg = p.dbg + 16 + (inst % 8) * 16;
      s_dbg[0] = gtimer();
    }
  }
#define DBG(k) do { if (s_dbg) s_dbg[k] = gtimer(); } while (0)
#else
#define DBG(k) do { } while (0)
#endif

  pdl_launch();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB);
    if (p.nterms > 1) tma_prefetch_desc(&tmA_lo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EPI_WARPS * 32);
    }
    mbar_init(bstat_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) DBG(1);

  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_n;
  const int total_tiles = tiles_m * p.tiles_co;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // Convergent warp: all lanes run the (uniform) loops and barrier waits, one elected lane issues.  Inside an
    // `if (lane == 0)` region ptxas wraps every UTMALDG / UTCHMMA in an ELECT / R2UR retry loop (see umma_bf16_pred).
    const uint32_t leader = elect_one() ? 1u : 0u;
    if ((int)blockIdx.x < total_tiles) {
      // weights of the first ring pass of the first tile: static data, fetched before the grid dependency resolves
      const int npre = kiters < STAGES ? kiters : STAGES;
      {
        const int co_t = blockIdx.x % p.tiles_co;
        if (p.b_stat) {
          mbar_expect_tx_pred(bstat_bar, kiters * B_STAGE_BYTES, leader);
          for (int it = 0; it < kiters; ++it)
            tma_load_2d_pred(sB + it * B_STAGE_BYTES, &tmB, bstat_bar, (it / p.kchunks) * p.Cin + (it % p.kchunks) * IG_BK, 0, leader);
          for (int it = 0; it < npre; ++it) mbar_expect_tx_pred(&full_bar[it], A_STAGE_BYTES, leader);
        } else
        for (int it = 0; it < npre; ++it) {
          const int kc = it % p.kchunks;
          const int t = (it / p.kchunks) % taps;
          const int term = it / (p.kchunks * taps);
          const int bko = (term == 1) ? p.b_lo_off : 0;
          mbar_expect_tx_pred(&full_bar[it], A_STAGE_BYTES + B_STAGE_BYTES, leader);
          tma_load_2d_pred(sB + it * B_STAGE_BYTES, &tmB, &full_bar[it], bko + t * p.Cin + kc * IG_BK, co_t * BN, leader);
        }
      }
      pdl_wait();
      if (leader) DBG(2);
      int git = 0;   // k-iterations issued so far by this CTA
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int co_t = tile % p.tiles_co;
        int mt = tile / p.tiles_co;
        const int tx = mt % p.tiles_x;
        mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int tn = mt / p.tiles_y;
        const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tn * p.bn;
        for (int term = 0; term < p.nterms; ++term) {
          const CUtensorMap* ma = (term == 2) ? &tmA_lo : &tmA_hi;
          const int bko = (term == 1) ? p.b_lo_off : 0;
          for (int t = 0; t < taps; ++t) {
            const int dy = t / p.taps_x - p.pad_y;
            const int dx = t % p.taps_x - p.pad_x;
            for (int kc = 0; kc < p.kchunks; ++kc, ++git) {
              const int stage = git % STAGES;
              if (git >= npre) {
                mbar_wait(&empty_bar[stage], ((git / STAGES) & 1) ^ 1);
                mbar_expect_tx_pred(&full_bar[stage], p.b_stat ? A_STAGE_BYTES : A_STAGE_BYTES + B_STAGE_BYTES, leader);
                if (!p.b_stat)
                  tma_load_2d_pred(sB + stage * B_STAGE_BYTES, &tmB, &full_bar[stage], bko + t * p.Cin + kc * IG_BK, co_t * BN, leader);
              }
              tma_load_4d_pred(sA + stage * A_STAGE_BYTES, ma, &full_bar[stage], kc * IG_BK, x0 + dx, y0 + dy, n0, leader);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t leader = elect_one() ? 1u : 0u;
    {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (p.b_stat && (int)blockIdx.x < total_tiles) mbar_wait(bstat_bar, 0);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (it == 0 && leader) DBG(3);
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * A_STAGE_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + (p.b_stat ? it : stage) * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < IG_BK / 16; ++k) {
            // +32 bytes (2 x 16 B units) per UMMA_K = 16 bf16 inside the 128 B swizzle atom
            umma_bf16_pred(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (it > 0 || k > 0) ? 1u : 0u, leader);
          }
          umma_commit_pred(&empty_bar[stage], leader);   // frees the smem stage when these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (leader) DBG(4);
        umma_commit_pred(&tfull_bar[acc], leader);       // accumulator ready for the epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps: 4 lane quarters x 2 column halves)
    const int e = warp - 2;
    const int quarter = warp & 3;           // TMEM lane window this warp may access
    const int half = e >> 2;                // which half of the tile's columns
    const int row = quarter * 32 + lane;
    const int r_tx = row % p.tw;
    const int r_ty = (row / p.tw) % p.th;
    const int r_bn = row / (p.tw * p.th);
    uint8_t* my_stage = sStage + e * UPW * UNIT_BYTES;
    long long* my_off = sOff + e * 64;      // [0..31] output pixel, [32..63] residual pixel (-1 = invalid row)
    const int u0 = (UNITS >= 2) ? half * (UNITS / 2) : 0;
    const int nu = (UNITS >= 2) ? UNITS / 2 : (half == 0 ? 1 : 0);
    const bool fast = (p.Cout & 7) == 0;    // 16-byte vector path needs Cout % 8 == 0
    const int t_r = lane >> 2, t_c = lane & 3;   // transposed mapping: 8 rows x 4 chunks per warp instruction
    int acc = 0;
    uint32_t acc_phase = 0;
    pdl_wait();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int co_t = tile % p.tiles_co;
      int mt = tile / p.tiles_co;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int tn = mt / p.tiles_y;
      const int x = tx * p.tw + r_tx, y = ty * p.th + r_ty, n = tn * p.bn + r_bn;
      const bool valid = (x < p.W) && (y < p.H) && (n < p.N);
      const size_t pix = ((size_t)n * p.H + y) * p.W + x;
      size_t rpix = pix;
      if (p.res_mode == RES_UP2) rpix = ((size_t)n * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1);
      __syncwarp();
      my_off[lane] = valid ? (long long)pix : -1;
      my_off[32 + lane] = valid ? (long long)rpix : -1;
      __syncwarp();
      // residual prefetch for every unit of this warp (in flight while the tile's MMAs are still running)
      if (p.res_mode != RES_NONE && fast) {
        if (p.tma_store) {                    // the staging buffers may still be the source of the previous tile's bulk stores
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        for (int j = 0; j < nu; ++j) {
          const int col0 = co_t * BN + (u0 + j) * 32;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + t_r;
            const long long off = my_off[32 + rr];
            const int col = col0 + t_c * 8;
            const bool ok = off >= 0 && col < p.Cout;
            const bf16* src = p.res.p + (ok ? (size_t)off * p.res.ld + col : 0);
            cp_async16(stage_ptr(my_stage + j * UNIT_BYTES, rr, t_c), src, ok);
          }
          cp_async_commit();
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (e == 0 && lane == 0) DBG(5);
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
      // Units are processed in PAIRS so that two independent instruction streams (TMEM load, staging round trip,
      // stores) overlap inside each latency-bound epilogue warp; the activation switch is hoisted out of the loops.
#pragma unroll 1
      for (int j = 0; j < nu; j += 2) {
        const bool two = (j + 1 < nu);
        uint32_t r[2][32];
        tmem_ld32(t_addr + (u0 + j) * 32, r[0]);
        if (two) tmem_ld32(t_addr + (u0 + j + 1) * 32, r[1]);
        tmem_ld_wait();
        const int colA = co_t * BN + (u0 + j) * 32;
        if (colA >= p.Cout) continue;        // warp-uniform (both units are beyond Cout)
        const int nun = (two && colA + 32 < p.Cout) ? 2 : 1;
        if (fast) {
          float v[2][32];
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < 32; ++i) v[q][i] = __uint_as_float(r[q][i]);
          // ---- bias (bias arrays are padded to a multiple of 32 floats)
          if (p.bias) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              if (q < nun) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + colA + q * 32 + i * 4));
                  v[q][4 * i] += b.x; v[q][4 * i + 1] += b.y; v[q][4 * i + 2] += b.z; v[q][4 * i + 3] += b.w;
                }
              }
            }
          }
          // ---- residual
          if (p.res_mode != RES_NONE) {
            cp_async_wait_pending(nu - j - nun);
            __syncwarp();
            if (e == 0 && lane == 0) DBG(6);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              if (q < nun) {
                uint8_t* buf = my_stage + (j + q) * UNIT_BYTES;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  float f[8];
                  unpack8(*stage_ptr(buf, lane, c), f);
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[q][c * 8 + i] += f[i];
                }
              }
            }
            __syncwarp();
            if (p.res.lo) {                  // split residual: low halves, synchronous transposed load
              for (int q = 0; q < nun; ++q) {
                uint8_t* buf = my_stage + (j + q) * UNIT_BYTES;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int rr = i * 8 + t_r;
                  const long long off = my_off[32 + rr];
                  const int col = colA + q * 32 + t_c * 8;
                  const bool ok = off >= 0 && col < p.Cout;
                  const bf16* src = p.res.p + p.res.lo + (ok ? (size_t)off * p.res.ld + col : 0);
                  cp_async16(stage_ptr(buf, rr, t_c), src, ok);
                }
              }
              cp_async_commit();
              cp_async_wait_pending(0);
              __syncwarp();
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                if (q < nun) {
                  uint8_t* buf = my_stage + (j + q) * UNIT_BYTES;
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    float f[8];
                    unpack8(*stage_ptr(buf, lane, c), f);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[q][c * 8 + i] += f[i];
                  }
                }
              }
              __syncwarp();
            }
          }
          // ---- activation (switch hoisted: one predictable branch per pair instead of three compares per element)
          if (p.act == ACT_RELU) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
              for (int i = 0; i < 32; ++i) v[q][i] = fmaxf(v[q][i], 0.f);
          } else if (p.act == ACT_LEAKY) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
              for (int i = 0; i < 32; ++i) v[q][i] = v[q][i] >= 0.f ? v[q][i] : 0.2f * v[q][i];
          } else if (p.act == ACT_RELU6) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
              for (int i = 0; i < 32; ++i) v[q][i] = fminf(fmaxf(v[q][i], 0.f), 6.f);
          }
          // ---- bf16 output, bulk tensor store: the pair's 32 rows x 64 channels are staged as 128-byte rows in the
          // SWIZZLE_128B pattern (thread = row: 16-byte chunk c at c ^ (row & 7), conflict-free) and ONE elected lane hands
          // them to the TMA unit, which writes whole lines and clips rows / channels outside the tensor.  The LSU version below
          // issues 8 x 64 B per warp instruction and, measured on C2, does not overlap with the rest of the epilogue: removing
          // the stores took the encoder from 9.6 to 7.8 ms although the kernels are not near the HBM write limit.
          if (p.out.p && p.tma_store) {
            uint8_t* buf = my_stage + j * UNIT_BYTES;          // 4 KB, 1 KB aligned (j is even)
            // this buffer's previous store (one tile ago) has been read.  (Allowing one group in flight for BN = 256 - two
            // pairs per warp - is wrong when a warp's second pair lies beyond Cout and is skipped: the one pending group is
            // then this very buffer's.  The TMA unit reads 4 KB of shared memory in well under a microsecond.)
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              if (q < nun) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  *reinterpret_cast<uint4*>(buf + lane * 128 + (((q * 4 + c) ^ (lane & 7)) << 4)) = pack8(v[q] + c * 8);
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              const int row0 = quarter * 32;
              const int x0 = tx * p.tw + row0 % p.tw, y0 = ty * p.th + (row0 / p.tw) % p.th, n0 = tn * p.bn + row0 / (p.tw * p.th);
              asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmA_lo),
                           "r"(smem_u32(buf)), "r"(colA), "r"(x0), "r"(y0), "r"(n0)
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          } else
          // ---- bf16 output through the staging buffers, transposed 16-byte stores
          if (p.out.p) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              if (q < nun) {
                uint8_t* buf = my_stage + (j + q) * UNIT_BYTES;
#pragma unroll
                for (int c = 0; c < 4; ++c) *stage_ptr(buf, lane, c) = pack8(v[q] + c * 8);
              }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = i * 8 + t_r;
              const long long off = my_off[rr];
              bf16* dst = p.out.p + (size_t)(off < 0 ? 0 : off) * p.out.ld + colA + t_c * 8;
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                if (q < nun && off >= 0 && colA + q * 32 + t_c * 8 < p.Cout)
                  *reinterpret_cast<uint4*>(dst + q * 32) = *stage_ptr(my_stage + (j + q) * UNIT_BYTES, rr, t_c);
              }
            }
            __syncwarp();
            if (p.out.lo) {
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                if (q < nun) {
                  uint8_t* buf = my_stage + (j + q) * UNIT_BYTES;
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    float hf[8], lo[8];
                    unpack8(pack8(v[q] + c * 8), hf);
#pragma unroll
                    for (int i = 0; i < 8; ++i) lo[i] = v[q][c * 8 + i] - hf[i];
                    *stage_ptr(buf, lane, c) = pack8(lo);
                  }
                }
              }
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rr = i * 8 + t_r;
                const long long off = my_off[rr];
                bf16* dst = p.out.p + p.out.lo + (size_t)(off < 0 ? 0 : off) * p.out.ld + colA + t_c * 8;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  if (q < nun && off >= 0 && colA + q * 32 + t_c * 8 < p.Cout)
                    *reinterpret_cast<uint4*>(dst + q * 32) = *stage_ptr(my_stage + (j + q) * UNIT_BYTES, rr, t_c);
                }
              }
              __syncwarp();
            }
          }
          // ---- fp32 output: two 16-column halves (64 B per row each) through the same buffer
          if (p.out_f32) {
            for (int q = 0; q < nun; ++q) {
              uint8_t* buf = my_stage + (j + q) * UNIT_BYTES;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  *stage_ptr(buf, lane, c) = make_uint4(__float_as_uint(v[q][hh * 16 + c * 4]), __float_as_uint(v[q][hh * 16 + c * 4 + 1]),
                                                        __float_as_uint(v[q][hh * 16 + c * 4 + 2]), __float_as_uint(v[q][hh * 16 + c * 4 + 3]));
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int rr = i * 8 + t_r;
                  const long long off = my_off[rr];
                  const int col = colA + q * 32 + hh * 16 + t_c * 4;
                  if (off >= 0 && col < p.Cout)
                    *reinterpret_cast<uint4*>(p.out_f32 + (size_t)off * p.ld_f32 + col) = *stage_ptr(buf, rr, t_c);
                }
                __syncwarp();
              }
            }
          }
        } else if (valid) {
          // ---- generic scalar path (Cout not a multiple of 8, e.g. the 1-channel score map)
          for (int q = 0; q < nun; ++q) {
            const int col0 = colA + q * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (col0 + i >= p.Cout) break;
              float t = __uint_as_float(r[q][i]);
              if (p.bias) t += p.bias[col0 + i];
              if (p.res_mode != RES_NONE) t += ld_act(p.res, rpix, col0 + i);
              t = apply_act(t, p.act);
              if (p.out.p) st_act(p.out, pix, col0 + i, t);
              if (p.out_f32) p.out_f32[pix * (size_t)p.ld_f32 + col0 + i] = t;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (e == 0 && lane == 0) DBG(7);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  if (warp >= 2 && lane == 0 && p.tma_store) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores performed
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) DBG(8);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
#undef DBG
}

template <int BN>
static int launch_bn(const IgemmOp& op, cudaStream_t stream) {
  FPNMT_CUDA_OK(launch_k(igemm_kernel<BN>, dim3(op.grid), dim3(IG_THREADS), (size_t)op.smem_bytes, stream, op.tmA_hi,
                         op.tmA_lo, op.tmB, op.p));
  return 0;
}

int igemm_launch(const IgemmOp& op, cudaStream_t stream) {
  switch (op.BN) {
    case 32: return launch_bn<32>(op, stream);
    case 64: return launch_bn<64>(op, stream);
    case 128: return launch_bn<128>(op, stream);
    case 256: return launch_bn<256>(op, stream);
  }
  set_last_error("igemm_launch: unsupported BN");
  return 1;
}

int igemm_set_attributes() {
  FPNMT_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, IG_SMEM_MAX));
  FPNMT_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, IG_SMEM_MAX));
  FPNMT_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, IG_SMEM_MAX));
  FPNMT_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, IG_SMEM_MAX));
  return 0;
}

}  // namespace fpnmt
