// What limits the rate of TMA loads issued by ONE thread?  (one CTA per SM, L2-resident source, 16 KB ring stages)
//   A  1 box of 128 rows per iteration, mbarrier.try_wait (suspending) on the stage issued `nst` iterations ago
//   B  same, polling with mbarrier.test_wait (never suspends)
//   C  4 boxes of 32 rows per iteration on the same barrier (same bytes as A)
//   D  like A from TWO threads of the SAME warp (lanes 0 and 1), own rings
//   E  like A from two threads of DIFFERENT warps on the same SM sub-partition (warps 0 and 4)
//   F  like A from two threads of warps on different sub-partitions (warps 0 and 1)
//   G  like A from four warps (0..3)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "tmap.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
using namespace pg;
constexpr int STAGE = 16384;

__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__device__ __forceinline__ void stream(const CUtensorMap* tm128, const CUtensorMap* tm32, uint32_t sb, uint64_t* bars, int ring, int nst, int iters,
                                       int rb, bool poll, bool four) {
  for (int it = 0; it < iters + nst; ++it) {
    const int s = it % nst;
    const uint32_t bar = smem_u32(&bars[ring * 4 + s]);
    if (it >= nst) {
      const uint32_t par = ((it / nst) - 1) & 1;
      if (poll) { while (!test_wait(bar, par)) {} } else mbar_wait(bar, par);
    }
    if (it < iters) {
      const int kb = it % 32;
      mbar_expect_tx(bar, STAGE);
      const uint32_t dst = sb + (ring * 3 + s) * STAGE;
      if (four) {
        for (int q = 0; q < 4; ++q) tma_load_2d(dst + q * 4096, tm32, bar, kb * 64, rb * 128 + q * 32, kEvictNormal);
      } else {
        tma_load_2d(dst, tm128, bar, kb * 64, rb * 128, kEvictNormal);
      }
    }
  }
}

__global__ void __launch_bounds__(256) k(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm32, int mode, int nst, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[16];
  const uint32_t sb = smem_u32(smem);
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1); mbar_fence_init(); }
  __syncthreads();
  const int rb = blockIdx.x % 32;
  const int t = threadIdx.x;
  int ring = -1;
  if (mode <= 2) ring = t == 0 ? 0 : -1;
  else if (mode == 3) ring = t == 0 ? 0 : t == 1 ? 1 : -1;
  else if (mode == 4) ring = t == 0 ? 0 : t == 128 ? 1 : -1;
  else if (mode == 5) ring = t == 0 ? 0 : t == 32 ? 1 : -1;
  else if (mode == 6) ring = (t % 32 == 0 && t < 128) ? t / 32 : -1;
  if (ring >= 0) stream(&tm128, &tm32, sb, bars, ring, nst, iters, (rb + 5 * ring) % 32, mode == 1, mode == 2);
}

int main() {
  const long long rows = 4096, cols = 2048;
  char* base; CK(cudaMalloc(&base, rows * cols * 2)); CK(cudaMemset(base, 1, rows * cols * 2));
  CUtensorMap tm128, tm32;
  if (make_tmap_2d(&tm128, base, rows, cols, cols, 128) != PG_OK || make_tmap_2d(&tm32, base, rows, cols, cols, 32) != PG_OK) return 1;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * STAGE + 1024));
  const char* names[] = {"A one thread, try_wait", "B one thread, test_wait polling", "C one thread, 4 boxes x 32 rows", "D two lanes of one warp", "E warps 0 and 4 (same sub-partition)", "F warps 0 and 1", "G warps 0..3"};
  const int nthreads[] = {1, 1, 1, 2, 2, 2, 4};
  const int iters = 32 * 16;
  for (int nst : {1, 2, 3}) {
    for (int mode = 0; mode < 7; ++mode) {
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      k<<<148, 256, 12 * STAGE + 1024>>>(tm128, tm32, mode, nst, iters); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      k<<<148, 256, 12 * STAGE + 1024>>>(tm128, tm32, mode, nst, iters);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double bytes = double(STAGE) * iters * nthreads[mode];
      printf("stages %d  %-40s %7.1f GB/s per SM   %6.0f ns per iteration\n", nst, names[mode], bytes / (ms * 1e-3) / 1e9, ms * 1e6 / iters);
    }
  }
  return 0;
}
