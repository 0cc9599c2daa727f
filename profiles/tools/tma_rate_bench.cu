// Per-SM global->shared delivery rate on B200 for L2-resident data, by path:
//   0  TMA 2D tensor loads, box = 128 rows x 128 B (the decode GEMM's weight tile, 128B swizzle)
//   1  1D bulk copies (cp.async.bulk) of 16 KB contiguous
//   2  LSU cp.async.cg 16 B per thread (128 threads)
//   3  0 and 2 at the same time (separate rings): do the two paths add up, or share one port?
//   4  TMA 2D with box = 64 rows x 128 B (the activation tile)
//   5  TMA 2D 128x128B, the issuing thread alternates between TWO tensor-map objects (same tensor)
//   6  TMA 2D 128x128B from TWO issuing threads (warps 0 and 1), own ring and own tensor map each
//   7  TMA 2D box = 32 rows x 128 B (4 KB)
// One CTA per SM, `nst` 16 KB stages in flight per path, nothing consumes the data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I paligemma_multimodal_system_b200/csrc -I include -o build/tma_rate_bench profiles/tools/tma_rate_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "tmap.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
using namespace pg;

constexpr int STAGE = 16384;

__global__ void __launch_bounds__(256) rate_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm64,
                                                   const __grid_constant__ CUtensorMap tm128b, const __grid_constant__ CUtensorMap tm32,
                                                   const char* base, int mode, int nst, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[16];
  const uint32_t sb = smem_u32(smem);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int rb = blockIdx.x % 32;  // 32 row blocks of 128 rows x 2048 columns (512 KB each): 16 MB, L2 resident
  long long t0 = clock64();
  const bool tma = mode == 0 || mode == 1 || mode == 3 || mode == 4 || mode == 5 || mode == 6 || mode == 7;
  if (mode == 6 && threadIdx.x == 32) {  // second issuing thread: stages [nst, 2 nst), barriers [8, 8 + nst)
    for (int it = 0; it < iters + nst; ++it) {
      const int s = it % nst;
      const uint32_t bar = smem_u32(&bars[8 + s]);
      if (it >= nst) mbar_wait(bar, ((it / nst) - 1) & 1);
      if (it < iters) {
        mbar_expect_tx(bar, STAGE);
        tma_load_2d(sb + (6 + s) * STAGE, &tm128b, bar, (it % 32) * 64, ((rb + 7) % 32) * 128, kEvictNormal);
      }
    }
  }
  if (tma && threadIdx.x == 0) {
    for (int it = 0; it < iters + nst; ++it) {
      const int s = it % nst;
      const uint32_t bar = smem_u32(&bars[s]);
      if (it >= nst) mbar_wait(bar, ((it / nst) - 1) & 1);
      if (it < iters) {
        const int kb = it % 32;
        mbar_expect_tx(bar, mode == 4 ? STAGE / 2 : mode == 7 ? STAGE / 4 : STAGE);
        if (mode == 1) bulk_copy_g2s(sb + s * STAGE, base + (static_cast<long long>(rb) * 32 + kb) * STAGE, STAGE, bar);
        else if (mode == 4) tma_load_2d(sb + s * STAGE, &tm64, bar, kb * 64, rb * 128 + (it & 1) * 64, kEvictNormal);
        else if (mode == 7) tma_load_2d(sb + s * STAGE, &tm32, bar, kb * 64, rb * 128 + (it & 3) * 32, kEvictNormal);
        else if (mode == 5 && (it & 1)) tma_load_2d(sb + s * STAGE, &tm128b, bar, kb * 64, rb * 128, kEvictNormal);
        else tma_load_2d(sb + s * STAGE, &tm128, bar, kb * 64, rb * 128, kEvictNormal);
      }
    }
  }
  if ((mode == 2 || mode == 3) && threadIdx.x >= 128) {
    const int t = threadIdx.x - 128;
    const uint32_t ring = sb + 7 * STAGE;
    for (int it = 0; it < iters; ++it) {
      const int s = it % nst, kb = it % 32;
      // stage = 128 rows x 128 B: thread t copies row t (8 x 16 B)
      const char* src = base + (static_cast<long long>(rb) * 128 + t) * 4096 + kb * 128;
      const uint32_t dst = ring + s * STAGE + t * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + ((c ^ (t & 7)) << 4)), "l"(src + c * 16) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (nst == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else if (nst == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
      else asm volatile("cp.async.wait_group 5;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  const long long rows = 4096, cols = 2048;
  char* base;
  CK(cudaMalloc(&base, rows * cols * 2));
  CK(cudaMemset(base, 1, rows * cols * 2));
  CUtensorMap tm128, tm64, tm128b, tm32;
  if (make_tmap_2d(&tm128, base, rows, cols, cols, 128) != PG_OK || make_tmap_2d(&tm64, base, rows, cols, cols, 64) != PG_OK ||
      make_tmap_2d(&tm128b, base, rows, cols, cols, 128) != PG_OK || make_tmap_2d(&tm32, base, rows, cols, cols, 32) != PG_OK) { printf("tmap failed\n"); return 1; }
  long long* out;
  CK(cudaMalloc(&out, 148 * 8));
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 13 * STAGE + 1024));
  const char* names[] = {"TMA 2D 128x128B", "bulk 1D 16 KB", "cp.async 16 B x128 thr", "TMA 2D + cp.async", "TMA 2D 64x128B (8 KB)", "TMA 2D, 2 tensor maps", "TMA 2D, 2 threads x 2 maps", "TMA 2D 32x128B (4 KB)"};
  const int iters = 32 * 16;
  for (int grid : {148}) {
    for (int nst : {2, 4, 6}) {
      for (int mode = 0; mode < 8; ++mode) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        rate_kernel<<<grid, 256, 13 * STAGE + 1024>>>(tm128, tm64, tm128b, tm32, base, mode, nst, iters, out);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        rate_kernel<<<grid, 256, 13 * STAGE + 1024>>>(tm128, tm64, tm128b, tm32, base, mode, nst, iters, out);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> h(grid);
        CK(cudaMemcpy(h.data(), out, grid * 8, cudaMemcpyDeviceToHost));
        double avg = 0; for (auto v : h) avg += v; avg /= grid;
        const double per = (mode == 4 ? STAGE / 2 : mode == 7 ? STAGE / 4 : STAGE) * ((mode == 3 || mode == 6) ? 2.0 : 1.0);
        const double bytes = per * iters;
        printf("grid %3d stages %d  %-24s %6.1f B/clk per SM   %7.1f GB/s per SM   aggregate %6.2f TB/s\n", grid, nst, names[mode],
               bytes / avg, bytes / (ms * 1e-3) / 1e9, bytes * grid / (ms * 1e-3) / 1e12);
      }
    }
  }
  return 0;
}
