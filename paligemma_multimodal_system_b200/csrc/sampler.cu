// Sampling kernels: greedy argmax (inference.py:67-68) and temperature + top-p (inference.py:63-66,90-106).
//
// Top-p without a sort: the kept set of _sample_top_p is {i : sum of probs strictly greater than p_i <= top_p}, i.e. a
// threshold on the logit.  The threshold is found by a 3-level radix select (11+11+10 bits of the order-preserving
// integer image of the fp32 logit) whose histogram bins accumulate probability MASS; the token is then drawn by inverse
// CDF over the kept set in vocabulary order with a counter-based RNG.  One CTA per row; the row (1 MB at V = 257 216)
// stays L2 resident across the passes.
#include "common.cuh"
#include "paligemma_b200.h"

namespace pg {

PG_DEVINL uint32_t float_key(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Visits every element of a row once: f(value, index).  16-byte loads, 4 independent loads in flight per thread
// (the row is L2 resident; one load per thread at a time is latency bound).
template <typename F>
PG_DEVINL void for_each_elem(const float* __restrict__ row, int V, F f) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if ((V & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    const int n4 = V >> 2;
    int i = tid;
    for (; i + 3 * nt < n4; i += 4 * nt) {
      const float4 a = __ldg(r4 + i), b = __ldg(r4 + i + nt), c = __ldg(r4 + i + 2 * nt), d = __ldg(r4 + i + 3 * nt);
      f(a.x, 4 * i); f(a.y, 4 * i + 1); f(a.z, 4 * i + 2); f(a.w, 4 * i + 3);
      f(b.x, 4 * (i + nt)); f(b.y, 4 * (i + nt) + 1); f(b.z, 4 * (i + nt) + 2); f(b.w, 4 * (i + nt) + 3);
      f(c.x, 4 * (i + 2 * nt)); f(c.y, 4 * (i + 2 * nt) + 1); f(c.z, 4 * (i + 2 * nt) + 2); f(c.w, 4 * (i + 2 * nt) + 3);
      f(d.x, 4 * (i + 3 * nt)); f(d.y, 4 * (i + 3 * nt) + 1); f(d.z, 4 * (i + 3 * nt) + 2); f(d.w, 4 * (i + 3 * nt) + 3);
    }
    for (; i < n4; i += nt) {
      const float4 a = __ldg(r4 + i);
      f(a.x, 4 * i); f(a.y, 4 * i + 1); f(a.z, 4 * i + 2); f(a.w, 4 * i + 3);
    }
  } else {
    for (int i = tid; i < V; i += nt) f(row[i], i);
  }
}

__global__ void __launch_bounds__(1024) argmax_kernel(const float* __restrict__ logits, long long ld, int* __restrict__ out, int V) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const float* row = logits + blockIdx.x * ld;
  griddep_wait();
  if (threadIdx.x == 0) griddep_launch_dependents();
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for_each_elem(row, V, [&](float v, int i) {
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = sv[lane]; bi = si[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) out[blockIdx.x] = bi;
  }
}

PG_DEVINL uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct BlockRed {
  float f[33];
  int i[33];
};

PG_DEVINL float block_reduce_max(float v, BlockRed& r) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) r.f[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = r.f[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (lane == 0) r.f[32] = t;
  }
  __syncthreads();
  return r.f[32];
}
PG_DEVINL float block_reduce_sum(float v, BlockRed& r) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) r.f[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = r.f[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) r.f[32] = t;
  }
  __syncthreads();
  return r.f[32];
}

// One CTA (1024 threads) per row.
__global__ void __launch_bounds__(1024) sample_top_p_kernel(const float* __restrict__ logits, long long ld, int* __restrict__ out,
                                                            int* __restrict__ kept_count, int V, float inv_temp, float top_p,
                                                            unsigned long long seed, const int* __restrict__ step_ptr) {
  constexpr int NB = 2048;
  __shared__ float h_mass[NB];
  __shared__ int h_cnt[NB];
  __shared__ BlockRed red;
  __shared__ float w_tot[32];
  __shared__ int s_bin;
  __shared__ float s_above_mass;
  __shared__ int s_above_cnt;
  __shared__ float s_bin_mass;
  __shared__ int s_bin_cnt;
  __shared__ int s_target_warp;
  __shared__ float s_target_off;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = logits + blockIdx.x * ld;
  griddep_wait();
  if (threadIdx.x == 0) griddep_launch_dependents();

  // pass 0: max
  float mx = -INFINITY;
  for_each_elem(row, V, [&](float v, int) { mx = fmaxf(mx, v); });
  mx = block_reduce_max(mx, red);
  const float c = inv_temp * 1.4426950408889634f;  // exp((x - mx) * inv_temp) = exp2((x - mx) * c)

  // radix select on the logit key, bins carry probability mass (unnormalised, relative to the row max)
  uint32_t prefix = 0;        // key bits decided so far
  float above_mass = 0.f;     // mass of keys strictly above the current prefix range
  int above_cnt = 0;
  float Z = 0.f;
  const int shifts[3] = {21, 10, 0};
  const int widths[3] = {11, 11, 10};
  for (int level = 0; level < 3; ++level) {
    const int shift = shifts[level], nb = 1 << widths[level];
    for (int i = tid; i < NB; i += 1024) { h_mass[i] = 0.f; h_cnt[i] = 0; }
    __syncthreads();
    float zloc = 0.f;
    const uint32_t hi_mask = (level == 0) ? 0u : (0xFFFFFFFFu << (shift + widths[level]));
    for_each_elem(row, V, [&](float x, int) {
      const uint32_t key = float_key(x);
      if (level == 0) {
        const float w = exp2f((x - mx) * c);
        zloc += w;
        const int bin = key >> shift;
        atomicAdd(&h_mass[bin], w);
        atomicAdd(&h_cnt[bin], 1);
      } else if ((key & hi_mask) == prefix) {
        const int bin = (key >> shift) & (nb - 1);
        atomicAdd(&h_mass[bin], exp2f((x - mx) * c));
        atomicAdd(&h_cnt[bin], 1);
      }
    });
    if (level == 0) Z = block_reduce_sum(zloc, red);
    __syncthreads();
    const float thresh = top_p * Z;
    // descending scan: thread t owns bins (nb-1-2t, nb-2-2t); "before" = mass of all higher bins (+ above_mass)
    const int b0 = nb - 1 - 2 * tid, b1 = nb - 2 - 2 * tid;
    const float m0 = (b0 >= 0) ? h_mass[b0] : 0.f, m1 = (b1 >= 0) ? h_mass[b1] : 0.f;
    const int c0 = (b0 >= 0) ? h_cnt[b0] : 0, c1 = (b1 >= 0) ? h_cnt[b1] : 0;
    float ms = m0 + m1;
    int cs = c0 + c1;
    float ims = ms;  // inclusive scans across threads
    int ics = cs;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float tf = __shfl_up_sync(0xffffffffu, ims, o);
      const int ti = __shfl_up_sync(0xffffffffu, ics, o);
      if (lane >= o) { ims += tf; ics += ti; }
    }
    __syncthreads();
    if (lane == 31) { red.f[warp] = ims; red.i[warp] = ics; }
    if (tid == 0) s_bin = 0x7fffffff;
    __syncthreads();
    float offm = above_mass;
    int offc = above_cnt;
    for (int w = 0; w < warp; ++w) { offm += red.f[w]; offc += red.i[w]; }
    const float before0 = offm + ims - ms, before1 = before0 + m0;
    const int cbefore0 = offc + ics - cs, cbefore1 = cbefore0 + c0;
    // lowest non-empty bin whose "before" mass is still <= thresh
    if (b1 >= 0 && c1 > 0 && before1 <= thresh) atomicMin(&s_bin, b1);
    else if (b0 >= 0 && c0 > 0 && before0 <= thresh) atomicMin(&s_bin, b0);
    __syncthreads();
    int sel = s_bin;
    if (sel == 0x7fffffff) sel = nb - 1;  // cannot happen for level 0 (the max element always qualifies); defensive
    if (b0 == sel) { s_above_mass = before0; s_above_cnt = cbefore0; s_bin_mass = m0; s_bin_cnt = c0; }
    if (b1 == sel) { s_above_mass = before1; s_above_cnt = cbefore1; s_bin_mass = m1; s_bin_cnt = c1; }
    __syncthreads();
    above_mass = s_above_mass;
    above_cnt = s_above_cnt;
    prefix |= static_cast<uint32_t>(sel) << shift;
    __syncthreads();
  }
  // kept set = keys >= prefix ; kept mass = above_mass + mass of the threshold value's ties
  const float kept_mass = above_mass + s_bin_mass;
  if (tid == 0 && kept_count) kept_count[blockIdx.x] = above_cnt + s_bin_cnt;

  // draw u in (0,1) and walk the kept set in vocabulary order
  const int step = step_ptr ? *step_ptr : 0;
  const uint64_t rnd = splitmix64(seed ^ splitmix64((static_cast<uint64_t>(step) << 32) | blockIdx.x));
  const float u = (static_cast<float>(rnd >> 40) + 0.5f) * (1.0f / 16777216.0f);
  const float target = u * kept_mass;

  // each warp owns a contiguous range (multiple of 128 elements: one float4 per lane per iteration)
  const bool vec = ((V & 3) == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  const int per_warp = ((V + 31) / 32 + 127) / 128 * 128;
  const int lo = warp * per_warp, hi = min(V, lo + per_warp);
  auto load4 = [&](int i, float (&x)[4]) {  // elements i..i+3 (out of range -> -inf: never kept, zero weight)
    if (vec && i + 3 < hi) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(row + i));
      x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = (i + j < hi) ? row[i + j] : -INFINITY;
    }
  };
  float wsum = 0.f;
  for (int base = lo; base < hi; base += 512) {  // 4 float4 loads in flight per lane
    float x[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) load4(base + u * 128 + lane * 4, x[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (float_key(x[u][j]) >= prefix && x[u][j] != -INFINITY) wsum += exp2f((x[u][j] - mx) * c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  if (lane == 0) w_tot[warp] = wsum;
  __syncthreads();
  if (tid == 0) {
    float acc = 0.f;
    int tw = -1;
    float off = 0.f;
    int last_nonempty = 0;
    float last_off = 0.f;
    for (int w = 0; w < 32; ++w) {
      if (w_tot[w] > 0.f) { last_nonempty = w; last_off = acc; }
      if (tw < 0 && w_tot[w] > 0.f && target < acc + w_tot[w]) { tw = w; off = acc; }
      acc += w_tot[w];
    }
    if (tw < 0) { tw = last_nonempty; off = last_off; }  // rounding: target fell past the end
    s_target_warp = tw;
    s_target_off = off;
  }
  __syncthreads();
  if (warp == s_target_warp) {
    float acc = s_target_off;
    int found = -1, last_kept = -1;
    for (int base = lo; base < hi && found < 0; base += 128) {
      const int i0 = base + lane * 4;
      float x[4], w[4];
      load4(i0, x);
      float lsum = 0.f;
      bool any_kept = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool kept = float_key(x[j]) >= prefix && x[j] != -INFINITY;
        w[j] = kept ? exp2f((x[j] - mx) * c) : 0.f;
        any_kept |= kept;
        lsum += w[j];
      }
      float inc = lsum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const bool hit = any_kept && (target < acc + inc);
      const uint32_t hits = __ballot_sync(0xffffffffu, hit);
      const uint32_t keeps = __ballot_sync(0xffffffffu, any_kept);
      if (keeps) {  // remember the last kept element seen so far (fallback for rounding at the very end)
        const int kl = 31 - __clz(keeps);
        int lk = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (w[j] > 0.f || (float_key(x[j]) >= prefix && x[j] != -INFINITY)) lk = i0 + j;
        last_kept = __shfl_sync(0xffffffffu, lk, kl);
      }
      if (hits) {
        const int hl = __ffs(hits) - 1;
        // inside the hitting lane: walk its 4 elements
        float a = acc + inc - lsum;
        int f = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (f < 0 && w[j] > 0.f) {
            a += w[j];
            if (target < a) f = i0 + j;
          }
        }
        if (f < 0) {  // rounding inside the lane: take its last kept element
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (w[j] > 0.f) f = i0 + j;
        }
        found = __shfl_sync(0xffffffffu, f, hl);
      }
      acc += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (found < 0) found = last_kept;
    if (lane == 0) out[blockIdx.x] = found;
  }
}

}  // namespace pg

using namespace pg;

extern "C" int pg_argmax(const float* logits, long long ld, int* out, int B, int V, void* stream) {
  if (B <= 0 || V <= 0) return PG_ERR_ARG;
  return launch_kernel(argmax_kernel, dim3(B), dim3(1024), 0, reinterpret_cast<cudaStream_t>(stream), logits, ld, out, V) == cudaSuccess
             ? PG_OK : PG_ERR_CUDA;
}

extern "C" int pg_sample_top_p(const float* logits, long long ld, int* out, int* kept_count, int B, int V,
                               float inv_temperature, float top_p, unsigned long long seed, const int* step_ptr,
                               void* stream) {
  if (B <= 0 || V <= 0 || !(inv_temperature > 0.f) || !(top_p >= 0.f)) return PG_ERR_ARG;
  return launch_kernel(sample_top_p_kernel, dim3(B), dim3(1024), 0, reinterpret_cast<cudaStream_t>(stream), logits, ld, out,
                       kept_count, V, inv_temperature, top_p, seed, step_ptr) == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}
