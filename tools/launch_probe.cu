// Micro-probe: how many dependent-chain kernels per microsecond can one B200 dispatch when S streams each replay a CUDA graph
// holding a chain of K small kernels?  The decode step of this repo is 38 dependent kernels per step and up to 8 such chains run
// concurrently (lanes); decode-only throughput saturates near 130 us per step with 8 lanes, i.e. ~3.4 us per kernel launch in
// aggregate - this probe separates a launch / dispatch limit from SM residency by using kernels that do (almost) nothing.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 launch_probe.cu -o launch_probe
// Variants: grid size (CTAs), dynamic shared memory per CTA, programmatic dependent launch edges, cluster launch attribute,
//           body duration (ns of spinning per CTA).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define OK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__global__ void k_small(int spin_ns, int pdl, int* sink) {
  extern __shared__ unsigned char smem[];
  if (pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  if (spin_ns > 0) {
    long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < spin_ns);
  }
  if (sink && threadIdx.x == 0 && blockIdx.x == 0x7fffffff) *sink = smem[0];
}

struct Variant {
  const char* name;
  int ctas, threads, smem, pdl, cluster, spin_ns;
};

static float run(const Variant& v, int streams, int chain, int reps) {
  std::vector<cudaStream_t> st(streams);
  std::vector<cudaGraphExec_t> ex(streams);
  OK(cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  for (int s = 0; s < streams; ++s) {
    OK(cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking));
    cudaGraph_t g;
    OK(cudaStreamBeginCapture(st[s], cudaStreamCaptureModeThreadLocal));
    for (int k = 0; k < chain; ++k) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(v.ctas);
      cfg.blockDim = dim3(v.threads);
      cfg.dynamicSmemBytes = v.smem;
      cfg.stream = st[s];
      cudaLaunchAttribute at[2];
      int n = 0;
      if (v.pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
      }
      if (v.cluster > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = v.cluster;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        ++n;
      }
      cfg.attrs = at;
      cfg.numAttrs = n;
      OK(cudaLaunchKernelEx(&cfg, k_small, v.spin_ns, v.pdl, (int*)nullptr));
    }
    OK(cudaStreamEndCapture(st[s], &g));
    OK(cudaGraphInstantiate(&ex[s], g, 0));
    OK(cudaGraphDestroy(g));
  }
  for (int s = 0; s < streams; ++s) OK(cudaGraphLaunch(ex[s], st[s]));   // warm-up
  OK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  OK(cudaEventCreate(&e0));
  OK(cudaEventCreate(&e1));
  std::vector<cudaEvent_t> done(streams);
  for (int s = 0; s < streams; ++s) OK(cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming));
  cudaStream_t main_s;
  OK(cudaStreamCreateWithFlags(&main_s, cudaStreamNonBlocking));
  OK(cudaEventRecord(e0, main_s));
  for (int s = 0; s < streams; ++s) OK(cudaStreamWaitEvent(st[s], e0, 0));
  for (int r = 0; r < reps; ++r)
    for (int s = 0; s < streams; ++s) OK(cudaGraphLaunch(ex[s], st[s]));
  for (int s = 0; s < streams; ++s) {
    OK(cudaEventRecord(done[s], st[s]));
    OK(cudaStreamWaitEvent(main_s, done[s], 0));
  }
  OK(cudaEventRecord(e1, main_s));
  OK(cudaDeviceSynchronize());
  float ms = 0.f;
  OK(cudaEventElapsedTime(&ms, e0, e1));
  for (int s = 0; s < streams; ++s) {
    OK(cudaGraphExecDestroy(ex[s]));
    OK(cudaStreamDestroy(st[s]));
    OK(cudaEventDestroy(done[s]));
  }
  OK(cudaStreamDestroy(main_s));
  return ms * 1e3f / ((float)reps * chain);   // us per chain link (all streams advance one link in this time)
}

int main() {
  const Variant vs[] = {
      {"1 CTA x 32 thr, no smem", 1, 32, 0, 0, 1, 0},
      {"48 CTAs x 192 thr, 100 KB smem", 48, 192, 100 * 1024, 0, 1, 0},
      {"48 CTAs x 192 thr, 100 KB smem, PDL", 48, 192, 100 * 1024, 1, 1, 0},
      {"16 CTAs x 192 thr, 100 KB smem, PDL, cluster 4", 16, 192, 100 * 1024, 1, 4, 0},
      {"64 CTAs x 192 thr, 218 KB smem, PDL", 64, 192, 218 * 1024, 1, 1, 0},
      {"48 CTAs x 192 thr, 100 KB smem, PDL, 4 us body", 48, 192, 100 * 1024, 1, 1, 4000},
      {"296 CTAs x 448 thr, 112 KB smem, PDL, 4 us body", 296, 448, 112 * 1024, 1, 1, 4000},
  };
  const int chain = 152, reps = 20;   // 152 = 4 decode steps of 38 kernels
  printf("us per chain link (one kernel of every stream's chain), graph replay, chain of %d kernels\n", chain);
  printf("%-52s %8s %8s %8s %8s   aggregate kernels/us at 8 streams\n", "variant", "1 str", "2 str", "4 str", "8 str");
  for (const Variant& v : vs) {
    float r[4];
    const int ss[4] = {1, 2, 4, 8};
    for (int i = 0; i < 4; ++i) r[i] = run(v, ss[i], chain, reps);
    printf("%-52s %8.2f %8.2f %8.2f %8.2f   %.2f\n", v.name, r[0], r[1], r[2], r[3], 8.f / r[3]);
  }
  return 0;
}
