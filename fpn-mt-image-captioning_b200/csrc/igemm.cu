// tcgen05 implicit-GEMM kernel (see igemm.cuh).  Warp-specialised, persistent:
//   warp 0      : TMA producer  (one elected lane)  global -> smem ring (A box + B box per stage)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane), tcgen05.mma 128 x BN x 16, fp32 in TMEM
//   warps 2..5  : epilogue, TMEM -> registers (tcgen05.ld 32x32b.x32) -> bias/residual/activation -> global
// Two TMEM accumulator buffers (2*BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#include "igemm.cuh"

namespace fpnmt {

constexpr int A_STAGE_BYTES = IG_BM * IG_BK * 2;   // 16 KB

__host__ __device__ constexpr int ig_stages(int BN) { return BN == 256 ? 4 : (BN == 128 ? 6 : 8); }
__host__ __device__ constexpr int ig_b_bytes(int BN) { return BN * IG_BK * 2; }

int igemm_stages(int BN) { return ig_stages(BN); }
size_t igemm_smem_bytes(int BN) {
  return (size_t)ig_stages(BN) * (A_STAGE_BYTES + ig_b_bytes(BN)) + 1024 /*align slack*/ + 256 /*barriers*/;
}

template <int BN>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
             const __grid_constant__ CUtensorMap tmB, const IgemmParams p) {
  constexpr int STAGES = ig_stages(BN);
  constexpr int B_STAGE_BYTES = ig_b_bytes(BN);
  constexpr int TMEM_COLS = 2 * BN;
  constexpr uint32_t IDESC = umma_idesc_bf16(IG_BM, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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
      mbar_init(&tempty_bar[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_n;
  const int total_tiles = tiles_m * p.tiles_co;
  const int taps = p.taps_y * p.taps_x;
  const int kiters = p.nterms * taps * p.kchunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
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
            for (int kc = 0; kc < p.kchunks; ++kc) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
              tma_load_4d(sA + stage * A_STAGE_BYTES, ma, &full_bar[stage], kc * IG_BK, x0 + dx, y0 + dy, n0);
              tma_load_2d(sB + stage * B_STAGE_BYTES, &tmB, &full_bar[stage], bko + t * p.Cin + kc * IG_BK,
                          co_t * BN);
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * A_STAGE_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < IG_BK / 16; ++k) {
            // +32 bytes (2 x 16 B units) per UMMA_K = 16 bf16 inside the 128 B swizzle atom
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);   // frees the smem stage when these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);       // accumulator ready for the epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps, 128 rows)
    const int quarter = warp & 3;           // TMEM lane window this warp may access
    const int row = quarter * 32 + lane;
    const int r_tx = row % p.tw;
    const int r_ty = (row / p.tw) % p.th;
    const int r_bn = row / (p.tw * p.th);
    int acc = 0;
    uint32_t acc_phase = 0;
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

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(t_addr + ch * 32, r);
        tmem_ld_wait();
        const int col0 = co_t * BN + ch * 32;
        if (valid && col0 < p.Cout) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int c = col0 + g * 8;
            if (c >= p.Cout) break;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
            if (c + 8 <= p.Cout) {
              if (p.bias) {
                const float4 b0 = *reinterpret_cast<const float4*>(p.bias + c);
                const float4 b1 = *reinterpret_cast<const float4*>(p.bias + c + 4);
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              if (p.res_mode != RES_NONE) {
                float rr[8];
                ld_act8(p.res, rpix, c, rr);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += rr[i];
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = apply_act(v[i], p.act);
              if (p.out.p) st_act8(p.out, pix, c, v);
              if (p.out_f32) {
                float* o = p.out_f32 + pix * (size_t)p.ld_f32 + c;
                *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
              }
            } else {
              for (int i = 0; i < 8 && c + i < p.Cout; ++i) {   // ragged tail of the channel range
                float t = v[i];
                if (p.bias) t += p.bias[c + i];
                if (p.res_mode != RES_NONE) t += ld_act(p.res, rpix, c + i);
                t = apply_act(t, p.act);
                if (p.out.p) st_act(p.out, pix, c + i, t);
                if (p.out_f32) p.out_f32[pix * (size_t)p.ld_f32 + c + i] = t;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int BN>
static int launch_bn(const IgemmOp& op, cudaStream_t stream) {
  igemm_kernel<BN><<<op.grid, IG_THREADS, igemm_smem_bytes(BN), stream>>>(op.tmA_hi, op.tmA_lo, op.tmB, op.p);
  FPNMT_CUDA_OK(cudaGetLastError());
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
  FPNMT_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)igemm_smem_bytes(32)));
  FPNMT_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)igemm_smem_bytes(64)));
  FPNMT_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)igemm_smem_bytes(128)));
  FPNMT_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)igemm_smem_bytes(256)));
  return 0;
}

}  // namespace fpnmt
