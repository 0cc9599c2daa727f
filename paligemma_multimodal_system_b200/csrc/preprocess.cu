// Image path of PaliGemmaProcessor on the GPU (processing_paligemma.py:13-73): PIL's 8-bit bicubic resize
// (image.resize((S, S), BICUBIC), :17-19), then * 1/255, (x - 0.5) / 0.5 and HWC -> CHW (:21-33, 52-71), bit-exact with the
// reference's CPU path.  Pillow's resample is two separable passes in 22-bit fixed point with a uint8 intermediate
// (Resample.c: ImagingResampleHorizontal_8bpc / Vertical_8bpc); the per-output-pixel windows and integer weights are
// computed by the host exactly as precompute_coeffs / normalize_coeffs_8bpc do (image_preprocess.py) and uploaded once
// per (input size, output size).  The value map uint8 -> float32 of rescale + normalise is a 256-entry table built by the
// host with the reference's own numpy expression, so the float rounding is the reference's by construction.
#include "common.cuh"
#include "paligemma_b200.h"

namespace pg {

constexpr int kPrecisionBits = 32 - 8 - 2;

PG_DEVINL int clip8(int v) {
  v >>= kPrecisionBits;  // arithmetic shift, as Pillow's clip8 lookup
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// src uint8 [H, W, 3] -> dst uint8 [H, S, 3]; one thread per (row, output column)
__global__ void __launch_bounds__(256) resample_h_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W,
                                                            int S, const int* __restrict__ kk, const int* __restrict__ bounds,
                                                            int ksize) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(H) * S) return;
  const int xx = static_cast<int>(idx % S), y = static_cast<int>(idx / S);
  const int x0 = bounds[2 * xx], n = bounds[2 * xx + 1];
  const int* __restrict__ k = kk + static_cast<long long>(xx) * ksize;
  const uint8_t* __restrict__ p = src + (static_cast<long long>(y) * W + x0) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < n; ++x) {
    const int w = __ldg(k + x);
    s0 += p[3 * x] * w;
    s1 += p[3 * x + 1] * w;
    s2 += p[3 * x + 2] * w;
  }
  uint8_t* o = dst + idx * 3;
  o[0] = static_cast<uint8_t>(clip8(s0));
  o[1] = static_cast<uint8_t>(clip8(s1));
  o[2] = static_cast<uint8_t>(clip8(s2));
}

// src uint8 [H, Wd, 3] -> out fp32 [3, S, Wd] = lut[resampled uint8]; one thread per (output row, column)
__global__ void __launch_bounds__(256) resample_v_u8_norm_kernel(const uint8_t* __restrict__ src, float* __restrict__ out, int H,
                                                                 int Wd, int S, const int* __restrict__ kk,
                                                                 const int* __restrict__ bounds, int ksize,
                                                                 const float* __restrict__ lut) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(S) * Wd) return;
  const int xx = static_cast<int>(idx % Wd), yy = static_cast<int>(idx / Wd);
  const int y0 = bounds[2 * yy], n = bounds[2 * yy + 1];
  const int* __restrict__ k = kk + static_cast<long long>(yy) * ksize;
  const uint8_t* __restrict__ p = src + (static_cast<long long>(y0) * Wd + xx) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int y = 0; y < n; ++y) {
    const int w = __ldg(k + y);
    const uint8_t* q = p + static_cast<long long>(y) * Wd * 3;
    s0 += q[0] * w;
    s1 += q[1] * w;
    s2 += q[2] * w;
  }
  const long long plane = static_cast<long long>(S) * Wd;
  out[idx] = __ldg(lut + clip8(s0));
  out[plane + idx] = __ldg(lut + clip8(s1));
  out[2 * plane + idx] = __ldg(lut + clip8(s2));
}

}  // namespace pg

using namespace pg;

extern "C" int pg_resample_h_u8(const void* src, void* dst, int H, int W, int S, const int* kk, const int* bounds, int ksize,
                                void* stream) {
  if (src == nullptr || dst == nullptr || kk == nullptr || bounds == nullptr || H <= 0 || W <= 0 || S <= 0 || ksize <= 0) return PG_ERR_ARG;
  const long long total = static_cast<long long>(H) * S;
  resample_h_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), H, W, S, kk, bounds, ksize);
  pg_count_launch(1);
  return cudaGetLastError() == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}

extern "C" int pg_resample_v_u8_norm(const void* src, float* out, int H, int Wd, int S, const int* kk, const int* bounds,
                                     int ksize, const float* lut, void* stream) {
  if (src == nullptr || out == nullptr || kk == nullptr || bounds == nullptr || lut == nullptr || H <= 0 || Wd <= 0 || S <= 0 || ksize <= 0)
    return PG_ERR_ARG;
  const long long total = static_cast<long long>(S) * Wd;
  resample_v_u8_norm_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src), out, H, Wd, S, kk, bounds, ksize, lut);
  pg_count_launch(1);
  return cudaGetLastError() == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}
