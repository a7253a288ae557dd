// Micro-probe: time N back-to-back tcgen05.mma (operands resident in shared memory) for several tile widths.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../fpn-mt-image-captioning_b200/csrc mma_probe.cu -o mma_probe
#include <cstdio>
#include "common.cuh"
using namespace fpnmt;
namespace fpnmt { void set_last_error(const std::string&) {} bool pdl_enabled() { return false; } }

__device__ __forceinline__ long long gt() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// mode bit0: distinct A slot per k-iter; bit1: commit to a side barrier every 4 MMAs; bit2: warps 2,3 spin on the final barrier;
// bit3: second issuer; bit4: CONVERGENT issue (the whole warp runs the loop, one elected lane issues through the predicated
// forms) instead of an `if (lane == 0)` region, where ptxas wraps every UTCHMMA in an ELECT / R2UR retry loop
template <int BN>
__global__ void __launch_bounds__(128, 1) probe(int n_mma, int distinct_a, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ uint64_t side[8];
  if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) mbar_init(&side[i], 1);
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * BN * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, BN);
  if ((distinct_a & 16) && warp == 1) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    long long g0 = gt();
    long long t0 = clock64();
    for (int it = 0; it < n_mma / 4; ++it) {
      const int ka = (distinct_a & 1) ? it % 4 : 0;
      const uint64_t ad = umma_desc_sw128(smem_u32(smem + ka * 16384));
      const uint64_t bd = umma_desc_sw128(smem_u32(smem + 4 * 16384 + (it % 4) * BN * 128));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_pred(tm, ad + 2 * k, bd + 2 * k, IDESC, (it > 0 || k > 0) ? 1u : 0u, leader);
      if (distinct_a & 2) umma_commit_pred(&side[it % 8], leader);
    }
    long long t1 = clock64();
    umma_commit_pred(&bar, leader);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    long long g1 = gt();
    if (leader) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
      out[2] = g1 - g0;
    }
  } else if (!(distinct_a & 16) && warp == 1 && lane == 0) {
    long long g0 = gt();
    long long t0 = clock64();
    for (int it = 0; it < n_mma / 4; ++it) {
      const int ka = (distinct_a & 1) ? it % 4 : 0;
      const uint64_t ad = umma_desc_sw128(smem_u32(smem + ka * 16384));
      const uint64_t bd = umma_desc_sw128(smem_u32(smem + 4 * 16384 + (it % 4) * BN * 128));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tm, ad + 2 * k, bd + 2 * k, IDESC, (it > 0 || k > 0) ? 1u : 0u);
      if (distinct_a & 2) umma_commit(&side[it % 8]);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    long long g1 = gt();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    out[2] = g1 - g0;
  } else if ((distinct_a & 8) && warp == 2 && lane == 0) {
    // second issuer on another scheduler, own accumulator columns, own completion barrier
    for (int it = 0; it < n_mma / 4; ++it) {
      const uint64_t ad = umma_desc_sw128(smem_u32(smem + ((it + 2) % 4) * 16384));
      const uint64_t bd = umma_desc_sw128(smem_u32(smem + 4 * 16384 + ((it + 2) % 4) * BN * 128));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tm + 256, ad + 2 * k, bd + 2 * k, IDESC, (it > 0 || k > 0) ? 1u : 0u);
    }
    umma_commit(&side[0]);
    mbar_wait(&side[0], 0);
    out[3] = clock64();
  } else if ((distinct_a & 4) && warp >= 2) {
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int BN>
void run(int n, int da) {
  long long* d; cudaMalloc(&d, 32);
  size_t smem = 4 * 16384 + 4 * BN * 128 + 1024;
  cudaFuncSetAttribute(probe<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) probe<BN><<<1, 128, smem>>>(n, da, d);
  long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("N=%3d n_mma=%4d mode=%d issue=%6lld clk total=%7lld clk = %6lld ns -> %.1f clk/MMA %.1f ns/MMA (%s)\n", BN, n, da, h[0], h[1], h[2],
         (double)h[1] / n, (double)h[2] / n, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  const int modes[4] = {1, 17, 3, 19};   // lane-0 vs convergent issue, without / with a commit every 4 MMAs
  for (int mi = 0; mi < 4; ++mi) {
    const int da = modes[mi];
    run<16>(256, da);
    run<32>(32, da); run<32>(256, da);
    run<64>(32, da); run<64>(256, da);
    run<128>(32, da); run<128>(256, da);
    run<256>(32, da); run<256>(256, da);
  }
  return 0;
}
