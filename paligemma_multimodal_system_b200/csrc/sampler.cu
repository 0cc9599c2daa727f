// Sampling kernels: greedy argmax (inference.py:67-68) and temperature + top-p (inference.py:63-66,90-106).
//
// Top-p without a sort: the kept set of _sample_top_p is {i : sum of probs strictly greater than p_i <= top_p}, i.e. a
// threshold on the logit.  The threshold is found by a 2-level histogram select (2048 x 2048 linear bins in logit space
// whose bins accumulate probability MASS); the token is then drawn by inverse CDF over the kept set in vocabulary order
// with a counter-based RNG.  One CTA per row; the row (1 MB at V = 257 216) stays L2 resident across the passes.
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

// One CTA (1024 threads) per row.  Four vectorised passes over the (L2-resident) row:
//   0. online max / sum-exp  ->  mx, Z
//   1. histogram of t = (mx - x) * inv_temp over [0, t_cut) in 2048 LINEAR bins carrying probability mass.  Elements with
//      t >= t_cut are provably outside the kept set (their total mass is < (1 - top_p) * Z / 128) and are skipped, which
//      also keeps shared-memory atomic contention low: the bins that matter are spread linearly in logit space.
//   2. the same inside the selected bin (2048 sub-bins): threshold resolved to t_cut / 2^22 (~5e-6 in logit units);
//      values closer than that are treated as ties and kept together.
//   3. inverse-CDF walk over the kept set {t < t*} in vocabulary order.
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
  __shared__ float s_tot_mass;
  __shared__ int s_bin_cnt;
  __shared__ int s_target_warp;
  __shared__ float s_target_off;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = logits + blockIdx.x * ld;
  griddep_wait();
  if (threadIdx.x == 0) griddep_launch_dependents();

  // pass 0: online softmax statistics
  const float c = inv_temp * 1.4426950408889634f;  // exp((x - mx) * inv_temp) = exp2((x - mx) * c)
  float m_loc = -INFINITY, z_loc = 0.f;
  for_each_elem(row, V, [&](float v, int) {
    if (v > m_loc) {
      z_loc = z_loc * exp2f((m_loc - v) * c) + 1.f;
      m_loc = v;
    } else {
      z_loc += exp2f((v - m_loc) * c);
    }
  });
  const float mx = block_reduce_max(m_loc, red);
  const float Z = block_reduce_sum(m_loc == -INFINITY ? 0.f : z_loc * exp2f((m_loc - mx) * c), red);
  const float thresh = top_p * Z;

  // elements with t >= t_cut cannot be kept: V * exp(-t_cut) <= (1 - top_p) * Z / 128
  float t_cut = 100.f;
  if (top_p < 1.f) t_cut = fminf(100.f, logf(static_cast<float>(V) * 128.f / ((1.f - top_p) * Z)));
  t_cut = fmaxf(t_cut, 1e-3f);

  // Every element gets a (coarse, fine) bin pair from FIXED float expressions, so the histogram passes and the final
  // kept test classify it identically:  b0 = floor(t * NB / t_cut),  b1 = clamp(floor((t - lo1) * NB / bw0)).
  const float inv_w0 = static_cast<float>(NB) / t_cut;
  const float bw0 = t_cut / static_cast<float>(NB);
  auto coarse_bin = [&](float x, float& t) -> int {  // -1: outside [0, t_cut) -> never kept
    t = (mx - x) * inv_temp;
    const float u = t * inv_w0;
    if (!(t < t_cut) || !(u < static_cast<float>(NB))) return -1;
    return min(NB - 1, static_cast<int>(u));
  };
  float lo1 = 0.f, inv_w1 = 0.f;
  auto fine_bin = [&](float t) -> int {
    const float u = (t - lo1) * inv_w1;
    return u < 0.f ? 0 : min(NB - 1, static_cast<int>(u));
  };
  int sel0 = 0, sel1 = 0;
  float above_mass = 0.f;  // mass of the bins before the selected one (all kept)
  int above_cnt = 0;
  const bool want_cnt = kept_count != nullptr;

  // Shared-memory atomics are the cost of a histogram pass (~1-2 cycles per element per SM), so the coarse level is
  // bracketed first: a 1/16 subsample locates the threshold bin to within +-D bins; the full pass then only bins
  // the elements inside the bracket (everything below it is summed in registers, everything above is ignored) and
  // VERIFIES that the bracket really contains the threshold -- otherwise it is redone over the whole range.
  auto scan_select = [&](float base_mass, int base_cnt, int& sel_out) {
    // ascending-t scan: thread owns bins (2 tid, 2 tid + 1); "before" = kept mass of all smaller t
    const int b0 = 2 * tid, b1 = 2 * tid + 1;
    const float m0 = h_mass[b0], m1 = h_mass[b1];
    const int c0 = h_cnt[b0], c1 = h_cnt[b1];
    const float ms = m0 + m1;
    const int cs = c0 + c1;
    float ims = ms;
    int ics = cs;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float tf = __shfl_up_sync(0xffffffffu, ims, o);
      const int ti = __shfl_up_sync(0xffffffffu, ics, o);
      if (lane >= o) { ims += tf; ics += ti; }
    }
    __syncthreads();
    if (lane == 31) { red.f[warp] = ims; red.i[warp] = ics; }
    if (tid == 0) s_bin = -1;
    __syncthreads();
    float offm = base_mass;
    int offc = base_cnt;
    for (int w = 0; w < warp; ++w) { offm += red.f[w]; offc += red.i[w]; }
    const float before0 = offm + ims - ms, before1 = before0 + m0;
    const int cbefore0 = offc + ics - cs, cbefore1 = cbefore0 + c0;
    // the LAST non-empty bin (largest t) whose preceding mass is still <= thresh holds the threshold
    if (m1 > 0.f && before1 <= thresh) atomicMax(&s_bin, b1);
    else if (m0 > 0.f && before0 <= thresh) atomicMax(&s_bin, b0);
    __syncthreads();
    const int sel = s_bin;
    if (b0 == sel) { s_above_mass = before0; s_above_cnt = cbefore0; s_bin_mass = m0; s_bin_cnt = c0; }
    if (b1 == sel) { s_above_mass = before1; s_above_cnt = cbefore1; s_bin_mass = m1; s_bin_cnt = c1; }
    if (tid == 1023) { s_tot_mass = before1 + m1; }  // mass of everything binned (+ base)
    __syncthreads();
    sel_out = sel;
  };
  auto clear_hist = [&]() {
    for (int i = tid; i < NB; i += 1024) { h_mass[i] = 0.f; h_cnt[i] = 0; }
    __syncthreads();
  };

  // ---- subsample (every 16th group of 4 elements) -> estimated threshold bin ----
  int bl = 0, bh = NB - 1;
  if (V >= 16384) {
    clear_hist();
    for (int i = tid * 64; i < V; i += 1024 * 64) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i + j < V) {
          float t;
          const float x = row[i + j];
          const int b0 = coarse_bin(x, t);
          if (b0 >= 0) atomicAdd(&h_mass[b0], 16.f * exp2f((x - mx) * c));
        }
      }
    }
    __syncthreads();
    int est;
    scan_select(0.f, 0, est);
    if (est >= 0) {
      const int D = max(8, static_cast<int>(0.25f * inv_w0) + 1);
      bl = max(0, est - D);
      bh = min(NB - 1, est + D);
    }
  }
  // ---- full coarse pass over the bracket (retry over the whole range if the bracket missed) ----
  for (int attempt = 0; attempt < 2; ++attempt) {
    clear_hist();
    float m_below = 0.f;
    int c_below = 0;
    for_each_elem(row, V, [&](float x, int) {
      float t;
      const int b0 = coarse_bin(x, t);
      if (b0 < 0 || b0 > bh) return;
      const float w = exp2f((x - mx) * c);
      if (b0 < bl) {
        m_below += w;
        ++c_below;
      } else {
        atomicAdd(&h_mass[b0], w);
        if (want_cnt) atomicAdd(&h_cnt[b0], 1);
      }
    });
    const float base_m = block_reduce_sum(m_below, red);
    const int base_c = static_cast<int>(block_reduce_sum(static_cast<float>(c_below), red) + 0.5f);
    __syncthreads();
    scan_select(base_m, base_c, sel0);
    const bool ok = (sel0 >= 0) && (bh == NB - 1 || s_tot_mass > thresh) && (bl == 0 || base_m <= thresh);
    __syncthreads();
    if (ok) break;
    bl = 0;
    bh = NB - 1;
    sel0 = 0;
  }
  if (sel0 < 0) sel0 = 0;
  above_mass = s_above_mass;
  above_cnt = s_above_cnt;
  lo1 = static_cast<float>(sel0) * bw0;
  inv_w1 = static_cast<float>(NB) / bw0;
  __syncthreads();
  // ---- fine level inside the selected coarse bin ----
  clear_hist();
  for_each_elem(row, V, [&](float x, int) {
    float t;
    if (coarse_bin(x, t) != sel0) return;
    const int bin = fine_bin(t);
    atomicAdd(&h_mass[bin], exp2f((x - mx) * c));
    if (want_cnt) atomicAdd(&h_cnt[bin], 1);
  });
  __syncthreads();
  scan_select(above_mass, above_cnt, sel1);
  if (sel1 < 0) sel1 = 0;
  above_mass = s_above_mass;
  above_cnt = s_above_cnt;
  // kept set: coarse bin < sel0, or coarse bin == sel0 and fine bin <= sel1 (the selected finest bin is kept whole)
  const float kept_mass = above_mass + s_bin_mass;
  if (tid == 0 && kept_count) kept_count[blockIdx.x] = above_cnt + s_bin_cnt;
  auto is_kept = [&](float x) -> bool {
    float t;
    const int b0 = coarse_bin(x, t);
    if (b0 < 0 || b0 > sel0) return false;
    if (b0 < sel0) return true;
    return fine_bin(t) <= sel1;
  };

  // draw u in (0,1) and walk the kept set in vocabulary order
  const int step = step_ptr ? *step_ptr : 0;
  const uint64_t rnd = splitmix64(seed ^ splitmix64((static_cast<uint64_t>(step) << 32) | blockIdx.x));
  const float u01 = (static_cast<float>(rnd >> 40) + 0.5f) * (1.0f / 16777216.0f);
  const float target = u01 * kept_mass;

  // each warp owns a contiguous range (multiple of 128 elements: one float4 per lane per iteration)
  const bool vec = ((V & 3) == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  const int per_warp = ((V + 31) / 32 + 127) / 128 * 128;
  const int lo_i = warp * per_warp, hi_i = min(V, lo_i + per_warp);
  auto load4 = [&](int i, float (&x)[4]) {  // elements i..i+3 (out of range -> -inf: never kept)
    if (vec && i + 3 < hi_i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(row + i));
      x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = (i + j < hi_i) ? row[i + j] : -INFINITY;
    }
  };
  float wsum = 0.f;
  for (int base = lo_i; base < hi_i; base += 512) {  // 4 float4 loads in flight per lane
    float x[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) load4(base + q * 128 + lane * 4, x[q]);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (is_kept(x[q][j])) wsum += exp2f((x[q][j] - mx) * c);
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
    for (int base = lo_i; base < hi_i && found < 0; base += 128) {
      const int i0 = base + lane * 4;
      float x[4], w[4];
      load4(i0, x);
      float lsum = 0.f;
      int lk = -1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool kept = is_kept(x[j]);
        w[j] = kept ? exp2f((x[j] - mx) * c) : -1.f;  // -1 marks "not kept" (a kept weight may underflow to 0)
        if (kept) { lk = i0 + j; lsum += w[j]; }
      }
      float inc = lsum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const bool hit = (lk >= 0) && (target < acc + inc);
      const uint32_t hits = __ballot_sync(0xffffffffu, hit);
      const uint32_t keeps = __ballot_sync(0xffffffffu, lk >= 0);
      if (keeps) last_kept = __shfl_sync(0xffffffffu, lk, 31 - __clz(keeps));
      if (hits) {
        const int hl = __ffs(hits) - 1;
        float a = acc + inc - lsum;  // mass before this lane's elements
        int f = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (f < 0 && w[j] >= 0.f) {
            a += w[j];
            if (target < a) f = i0 + j;
          }
        }
        if (f < 0) f = lk;  // rounding inside the lane
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
