// Cost of one global dependency between two kernels of a CUDA-graph chain on B200, three ways:
//   (a) plain stream order;
//   (b) programmatic dependent launch, consumer blocks in griddepcontrol.wait (what the decode chain does today);
//   (c) programmatic dependent launch WITHOUT griddepcontrol.wait: every producer CTA publishes "my stores are done" with a
//       release-add on a per-kernel counter, the consumer's CTAs poll it with ld.acquire.
// Each kernel: wait -> every thread reads 16 B the previous kernel wrote (one dependent L2 round trip) -> writes 16 B -> signal.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o boundary_bench boundary_bench.cu && ./boundary_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) link_kernel(const float4* in, float4* out, unsigned* counters, int k, unsigned expect, int work) {
  if (MODE == 1) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  if (MODE == 2) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (k > 0) {
      if (threadIdx.x == 0) {
        unsigned v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counters + (k - 1) * 32) : "memory");
        } while (v < expect);
      }
      __syncthreads();
    }
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float4 v = in[(i * 7 + 13) % (gridDim.x * blockDim.x)];
  for (int w = 0; w < work; ++w) v.x = v.x * 1.0001f + v.y;
  v.x += 1.f;
  out[i] = v;
  if (MODE == 2) {
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counters + k * 32) : "memory");
  }
}

template <int MODE>
static float run(int grid, int chain, int work, int reps) {
  float4 *a, *b;
  unsigned* counters;
  const size_t n = size_t(grid) * 256;
  CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16)); CK(cudaMalloc(&counters, chain * 128));
  CK(cudaMemset(a, 0, n * 16)); CK(cudaMemset(b, 0, n * 16));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  CK(cudaMemsetAsync(counters, 0, chain * 128, st));
  for (int k = 0; k < chain; ++k) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (MODE != 0 && k > 0) ? 1 : 0;
    const float4* in = (k & 1) ? b : a; float4* out = (k & 1) ? a : b;
    CK(cudaLaunchKernelEx(&cfg, link_kernel<MODE>, in, out, counters, k, unsigned(grid), work));
  }
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int t = 0; t < 3; ++t) {
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  // check: every element went through `chain` increments per replay
  std::vector<float4> h(n);
  CK(cudaMemcpy(h.data(), (chain & 1) ? b : a, n * 16, cudaMemcpyDeviceToHost));
  CK(cudaFree(a)); CK(cudaFree(b)); CK(cudaFree(counters));
  CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g)); CK(cudaStreamDestroy(st));
  return best * 1e3f / (reps * chain);
}

int main() {
  const int chain = 126;
  for (int work : {0, 2000}) {
    for (int grid : {64, 148, 296}) {
      float t0 = run<0>(grid, chain, work, 20), t1 = run<1>(grid, chain, work, 20), t2 = run<2>(grid, chain, work, 20);
      printf("grid %3d work %4d: stream order %6.2f us | PDL wait %6.2f us | PDL + release/acquire counter %6.2f us  per kernel\n", grid, work, t0, t1, t2);
    }
  }
  return 0;
}
