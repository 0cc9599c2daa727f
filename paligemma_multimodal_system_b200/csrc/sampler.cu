// Sampling kernels: greedy argmax (inference.py:67-68) and temperature + top-p (inference.py:63-66,90-106).
//
// Top-p without a sort: the kept set of _sample_top_p is {i : sum of probs strictly greater than p_i <= top_p}, i.e. a
// threshold on the logit.  The threshold is found by a 2-level histogram select (2048 x 2048 linear bins in logit space
// whose bins accumulate probability MASS); the token is then drawn by inverse CDF over the kept set in vocabulary order
// with a counter-based RNG.  One CTA per row; the row (1 MB at V = 257 216) stays L2 resident across the passes.
#include <cooperative_groups.h>

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

// four block-wide sums in one exchange (the verification pass of the rejection samplers carries four candidates)
PG_DEVINL void block_reduce_sum4(float (&v)[4], float (*buf)[4], float (&out)[4]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  __syncthreads();  // protect `buf` from its previous use
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) buf[warp][k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += buf[w][k];  // same order in every thread: identical results block-wide
    out[k] = t;
  }
}

// -------------------------------------------------------------------------------------------------------------------
// top-p sampler: one thread-block CLUSTER of R CTAs (1024 threads each) per row; each warp owns a contiguous segment of
// the row; cross-CTA reductions and histogram merges go through distributed shared memory.
//
//   P0  row max                                                       (no transcendental)
//   PS  1/16 subsample -> histogram of u = (mx - x) * inv_temp * NB / t_cut -> estimated threshold bin -> bracket [bl, bh]
//   P1  full pass: w = exp2(..) for every element; mass below / inside (shared-memory atomics, only the bracket) / above
//       the bracket -> exact Z, exact coarse bin sel0; the bracket is VERIFIED and the pass redone over all bins if the
//       estimate missed (shared-memory atomics cost ~1-2 cycles per element, hence the bracket)
//   P2  elements of the bracket: bins below sel0 are kept (segment sums), bin sel0 goes into a 2048-bin fine histogram and
//       a small candidate list -> fine bin sel1.  The threshold is thereby resolved to t_cut / 2^22 (~5e-6 in logit
//       units); values closer than that are ties and are kept together.
//   P3  inverse-CDF walk in vocabulary order: segment sums pick the segment, its warp re-reads ~V/(32R) elements.
// Every element is classified by ONE fixed float expression (u, and (u - sel0) * NB for the fine bin) in all passes.
// -------------------------------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;

__device__ int g_topp_retries = 0;
__device__ int g_topp_bracket = 0;
__device__ long long g_topp_trace[16];  // profiling: clock64 of CTA 0 / thread 0 after each phase of the last launch   // profiling: > 0 overrides the half-width (in coarse bins) of the estimated bracket  // profiling: how often the estimated bracket failed verification

struct TopPShared {
  float h_mass[2048];
  int h_cnt[2048];
  float m_sum[2048];      // merged (cluster-wide) histogram
  int c_sum[2048];
  float seg_kept[32];     // kept mass of each warp's segment
  int seg_cnt[32];
  float cta_val[4];       // per-CTA scalars published to the cluster: [0] max, [1..3] the sums that ride on merge_hist
  int cta_cnt[2];         // [0] below count
  float cand_w[1024];     // elements of coarse bin sel0: weight, fine bin, owning warp
  short cand_fine[1024];
  short cand_warp[1024];
  int n_cand;
  BlockRed red;
  int s_bin;
  float s_before_mass, s_bin_mass, s_tot_mass;
  int s_before_cnt, s_bin_cnt;
};

// visits the float4 groups of [lo, hi) (lo multiple of 4 elements) owned by one warp: f(value, index); 4 loads in flight
template <typename F>
PG_DEVINL void for_each_in_segment(const float* __restrict__ row, int lo, int hi, bool vec, int lane, F f) {
  if (vec) {
    int i = lo + lane * 4;
    for (; i + 3 * 128 + 3 < hi; i += 4 * 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(row + i));
      const float4 b = __ldg(reinterpret_cast<const float4*>(row + i + 128));
      const float4 c = __ldg(reinterpret_cast<const float4*>(row + i + 256));
      const float4 d = __ldg(reinterpret_cast<const float4*>(row + i + 384));
      f(a.x, i); f(a.y, i + 1); f(a.z, i + 2); f(a.w, i + 3);
      f(b.x, i + 128); f(b.y, i + 129); f(b.z, i + 130); f(b.w, i + 131);
      f(c.x, i + 256); f(c.y, i + 257); f(c.z, i + 258); f(c.w, i + 259);
      f(d.x, i + 384); f(d.y, i + 385); f(d.z, i + 386); f(d.w, i + 387);
    }
    for (; i + 3 < hi; i += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(row + i));
      f(a.x, i); f(a.y, i + 1); f(a.z, i + 2); f(a.w, i + 3);
    }
    for (int j = i; j < hi && j < i + 4; ++j) f(row[j], j);  // (hi is a multiple of 4 when vec: never taken)
  } else {
    for (int i = lo + lane; i < hi; i += 32) f(row[i], i);
  }
}

__global__ void __launch_bounds__(1024) sample_top_p_kernel(const float* __restrict__ logits, long long ld, int* __restrict__ out,
                                                            int* __restrict__ kept_count, int V, float inv_temp, float top_p,
                                                            unsigned long long seed, const int* __restrict__ step_ptr) {
  constexpr int NB = 2048;
  __shared__ TopPShared S;
  cg::cluster_group cluster = cg::this_cluster();
  const int R = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const int row_idx = blockIdx.x / R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ row = logits + row_idx * ld;
  griddep_wait();
  if (threadIdx.x == 0) griddep_launch_dependents();

  const bool vec = ((V & 3) == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  const int nseg = R * 32;
  const int seg_len = ((V + nseg - 1) / nseg + 127) / 128 * 128;
  const int seg = rank * 32 + warp;
  const int seg_lo = min(V, seg * seg_len), seg_hi = min(V, seg_lo + seg_len);
  const bool want_cnt = kept_count != nullptr;

  auto remote = [&](auto* ptr, int r) { return cluster.map_shared_rank(ptr, r); };
  auto cluster_sum_f = [&](float v, int slot) -> float {  // block reduce, publish, sum over the cluster in rank order
    const float t = block_reduce_sum(v, S.red);
    if (tid == 0) S.cta_val[slot] = t;
    cluster.sync();
    float acc = 0.f;
    for (int r = 0; r < R; ++r) acc += *remote(&S.cta_val[slot], r);
    cluster.sync();
    return acc;
  };

  const bool trc = blockIdx.x == 0 && tid == 0;
  if (trc) g_topp_trace[0] = clock64();
  // ---- P0: row max ----
  float m_loc = -INFINITY;
  for_each_in_segment(row, seg_lo, seg_hi, vec, lane, [&](float v, int) { m_loc = fmaxf(m_loc, v); });
  {
    const float t = block_reduce_max(m_loc, S.red);
    if (tid == 0) S.cta_val[0] = t;
    cluster.sync();
    float acc = -INFINITY;
    for (int r = 0; r < R; ++r) acc = fmaxf(acc, *remote(&S.cta_val[0], r));
    m_loc = acc;  // (cta_val[0] is never written again in this launch: no second barrier needed)
  }
  const float mx = m_loc;
  const float c = inv_temp * 1.4426950408889634f;  // w = exp((x - mx) * inv_temp) = exp2(x * c - mx * c)
  const float cm = mx * c;
  // elements with t = (mx - x) * inv_temp >= t_cut cannot be kept: V * exp(-t_cut) <= (1 - top_p) / 128 <= (1 - top_p) Z / 128
  float t_cut = 80.f;
  if (top_p < 1.f) t_cut = fminf(80.f, logf(static_cast<float>(V) * 128.f / (1.f - top_p)));
  t_cut = fmaxf(t_cut, 1e-3f);
  const float k0 = inv_temp * static_cast<float>(NB) / t_cut;  // u = (mx - x) * k0  in [0, NB) <=> t in [0, t_cut)
  const float c0 = mx * k0;
  // (clamped at 0: c0 = mx * k0 is rounded, so the fused multiply-add can come out a hair NEGATIVE for the row maximum
  //  itself, which would put it below bin 0 and leave every histogram empty)
  auto u_of = [&](float x) { return fmaxf(fmaf(-x, k0, c0), 0.f); };
  auto w_of = [&](float x) { return exp2f(fmaf(x, c, -cm)); };

  auto clear_hist = [&]() {
    for (int i = tid; i < NB; i += 1024) { S.h_mass[i] = 0.f; S.h_cnt[i] = 0; }
    if (tid < 32) { S.seg_kept[tid] = 0.f; S.seg_cnt[tid] = 0; }
    if (tid == 0) S.n_cand = 0;
    __syncthreads();
  };
  // merged histogram of the cluster (fixed rank order => bitwise identical in every CTA).  `nscal` per-CTA scalars
  // (block-reduced by the caller into S.cta_val[1..nscal]) ride on the same two barriers; their cluster sums land in sums[].
  float sums[3] = {0.f, 0.f, 0.f};
  auto merge_hist = [&](int nscal = 0) {
    cluster.sync();
    for (int k = 0; k < nscal; ++k) {
      float acc = 0.f;
      for (int r = 0; r < R; ++r) acc += remote(S.cta_val, r)[1 + k];
      sums[k] = acc;
    }
    for (int i = tid; i < NB; i += 1024) {
      float m = 0.f;
      int n = 0;
      for (int r = 0; r < R; ++r) { m += remote(S.h_mass, r)[i]; n += remote(S.h_cnt, r)[i]; }
      S.m_sum[i] = m;
      S.c_sum[i] = n;
    }
    cluster.sync();
  };
  // ascending-t scan of the merged histogram: the LAST non-empty bin whose preceding mass is <= thresh
  auto scan_select = [&](float base_mass, int base_cnt, float thresh) -> int {
    const int b0 = 2 * tid, b1 = 2 * tid + 1;
    const float m0 = S.m_sum[b0], m1 = S.m_sum[b1];
    const int n0 = S.c_sum[b0], n1 = S.c_sum[b1];
    const float ms = m0 + m1;
    const int cs = n0 + n1;
    float ims = ms;
    int ics = cs;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float tf = __shfl_up_sync(0xffffffffu, ims, o);
      const int ti = __shfl_up_sync(0xffffffffu, ics, o);
      if (lane >= o) { ims += tf; ics += ti; }
    }
    __syncthreads();
    if (lane == 31) { S.red.f[warp] = ims; S.red.i[warp] = ics; }
    if (tid == 0) S.s_bin = -1;
    __syncthreads();
    float offm = base_mass;
    int offc = base_cnt;
    for (int w = 0; w < warp; ++w) { offm += S.red.f[w]; offc += S.red.i[w]; }
    const float before0 = offm + ims - ms, before1 = before0 + m0;
    const int cb0 = offc + ics - cs, cb1 = cb0 + n0;
    if (m1 > 0.f && before1 <= thresh) atomicMax(&S.s_bin, b1);
    else if (m0 > 0.f && before0 <= thresh) atomicMax(&S.s_bin, b0);
    __syncthreads();
    const int sel = S.s_bin;
    if (b0 == sel) { S.s_before_mass = before0; S.s_before_cnt = cb0; S.s_bin_mass = m0; S.s_bin_cnt = n0; }
    if (b1 == sel) { S.s_before_mass = before1; S.s_before_cnt = cb1; S.s_bin_mass = m1; S.s_bin_cnt = n1; }
    if (tid == 1023) S.s_tot_mass = before1 + m1;
    __syncthreads();
    return sel;
  };

  if (trc) g_topp_trace[1] = clock64();
  // ---- PS: subsample -> bracket ----
  int bl = 0, bh = NB - 1;
  if (V >= 16384) {
    clear_hist();
    float zs = 0.f;
    for (int i = seg_lo + lane * 64; i < seg_hi; i += 32 * 64) {  // 4 consecutive elements out of every 64
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i + j < seg_hi) {
          const float x = row[i + j];
          const float u = u_of(x), w = 16.f * w_of(x);
          zs += w;
          if (u < static_cast<float>(NB)) atomicAdd(&S.h_mass[static_cast<int>(u)], w);
        }
      }
    }
    {
      const float t = block_reduce_sum(zs, S.red);
      if (tid == 0) S.cta_val[1] = t;
    }
    merge_hist(1);
    const float z_est = sums[0];
    const int est = scan_select(0.f, 0, top_p * z_est);
    if (est >= 0) {
      const int D = g_topp_bracket > 0 ? g_topp_bracket : 8;
      bl = max(0, est - D);
      bh = min(NB - 1, est + D);
    }
  }

  if (trc) g_topp_trace[2] = clock64();
  // ---- P1: exact Z, exact coarse bin (bracket verified, else redone over every bin) ----
  int sel0 = 0;
  float Z = 0.f, thresh = 0.f, base_mass = 0.f;
  int base_cnt = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    clear_hist();
    const float blf = static_cast<float>(bl), bhf = static_cast<float>(bh + 1);
    float m_below = 0.f, m_rest = 0.f;
    int n_below = 0;
    for_each_in_segment(row, seg_lo, seg_hi, vec, lane, [&](float x, int) {
      const float u = u_of(x), w = w_of(x);
      if (u < blf) {
        m_below += w;
        ++n_below;
      } else {
        m_rest += w;
        if (u < bhf) {
          const int bin = static_cast<int>(u);
          atomicAdd(&S.h_mass[bin], w);
          if (want_cnt) atomicAdd(&S.h_cnt[bin], 1);
        }
      }
    });
    // per-warp kept mass so far (everything below the bracket is kept if the bracket verifies)
    float wm = m_below;
    int wn = n_below;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wm += __shfl_xor_sync(0xffffffffu, wm, o);
      wn += __shfl_xor_sync(0xffffffffu, wn, o);
    }
    if (lane == 0) { S.seg_kept[warp] = wm; S.seg_cnt[warp] = wn; }
    {
      const float t1 = block_reduce_sum(m_below, S.red);
      const float t2 = block_reduce_sum(m_rest, S.red);
      const float t3 = block_reduce_sum(static_cast<float>(n_below), S.red);
      if (tid == 0) { S.cta_val[1] = t1; S.cta_val[2] = t2; S.cta_val[3] = t3; }
    }
    merge_hist(3);
    base_mass = sums[0];
    const float rest = sums[1];
    base_cnt = static_cast<int>(sums[2] + 0.5f);
    Z = base_mass + rest;
    thresh = top_p * Z;
    sel0 = scan_select(base_mass, base_cnt, thresh);
    const bool ok = (sel0 >= 0) && (bh == NB - 1 || S.s_tot_mass > thresh) && (bl == 0 || base_mass <= thresh);
    __syncthreads();
    if (ok) break;
    if (tid == 0 && rank == 0) atomicAdd(&g_topp_retries, 1);
    bl = 0;
    bh = NB - 1;
  }
  if (sel0 < 0) sel0 = 0;  // unreachable: the row maximum sits in bin 0 with nothing before it
  const float sel0f = static_cast<float>(sel0);
  float above_mass = S.s_before_mass;  // kept mass of every bin before sel0
  int above_cnt = S.s_before_cnt;
  __syncthreads();

  if (trc) g_topp_trace[3] = clock64();
  // ---- P2: bins [bl, sel0) of the bracket are kept; bin sel0 -> fine histogram + candidate list ----
  {
    const float blf = static_cast<float>(bl);
    for (int i = tid; i < NB; i += 1024) { S.h_mass[i] = 0.f; S.h_cnt[i] = 0; }
    __syncthreads();
    float m_in = 0.f;
    int n_in = 0;
    for_each_in_segment(row, seg_lo, seg_hi, vec, lane, [&](float x, int) {
      const float u = u_of(x);
      if (u < blf || !(u < sel0f + 1.f)) return;
      const float w = w_of(x);
      if (u < sel0f) {
        m_in += w;
        ++n_in;
      } else {
        const float u1 = (u - sel0f) * static_cast<float>(NB);
        const int fine = min(NB - 1, static_cast<int>(u1));
        atomicAdd(&S.h_mass[fine], w);
        if (want_cnt) atomicAdd(&S.h_cnt[fine], 1);
        const int slot = atomicAdd(&S.n_cand, 1);
        if (slot < 1024) { S.cand_w[slot] = w; S.cand_fine[slot] = static_cast<short>(fine); S.cand_warp[slot] = static_cast<short>(warp); }
      }
    });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m_in += __shfl_xor_sync(0xffffffffu, m_in, o);
      n_in += __shfl_xor_sync(0xffffffffu, n_in, o);
    }
    if (lane == 0) { S.seg_kept[warp] += m_in; S.seg_cnt[warp] += n_in; }
  }
  merge_hist();
  int sel1 = scan_select(above_mass, above_cnt, thresh);
  if (sel1 < 0) sel1 = 0;
  const int kept_total_cnt = S.s_before_cnt + S.s_bin_cnt;
  __syncthreads();
  auto is_kept = [&](float x, float& w) -> bool {
    const float u = u_of(x);
    if (!(u < sel0f + 1.f)) return false;
    w = w_of(x);
    if (u < sel0f) return true;
    return min(NB - 1, static_cast<int>((u - sel0f) * static_cast<float>(NB))) <= sel1;
  };
  // candidates of bin sel0 that made it (fine bin <= sel1) join their segment's kept mass
  if (S.n_cand <= 1024) {
    for (int i = tid; i < S.n_cand; i += 1024)
      if (S.cand_fine[i] <= sel1) atomicAdd(&S.seg_kept[S.cand_warp[i]], S.cand_w[i]);
  } else {  // pathological (thousands of near-identical logits at the threshold): recount bin sel0 from the row
    float m_c = 0.f;
    for_each_in_segment(row, seg_lo, seg_hi, vec, lane, [&](float x, int) {
      const float u = u_of(x);
      float w;
      if (!(u < sel0f) && is_kept(x, w)) m_c += w;
    });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m_c += __shfl_xor_sync(0xffffffffu, m_c, o);
    if (lane == 0) S.seg_kept[warp] += m_c;
  }
  if (tid == 0 && rank == 0 && want_cnt) kept_count[row_idx] = kept_total_cnt;

  if (trc) g_topp_trace[4] = clock64();
  // ---- P3: pick the segment, walk it ----
  cluster.sync();  // seg_kept of every CTA is final
  const int step = step_ptr ? *step_ptr : 0;
  const uint64_t rnd = splitmix64(seed ^ splitmix64((static_cast<uint64_t>(step) << 32) | static_cast<uint32_t>(row_idx)));
  const float u01 = (static_cast<float>(rnd >> 40) + 0.5f) * (1.0f / 16777216.0f);
  {
    // gather the (<= 128) segment sums into local shared memory, one thread scans them
    if (tid < nseg) S.m_sum[tid] = remote(S.seg_kept, tid >> 5)[tid & 31];
    __syncthreads();
    if (tid == 0) {
      float total = 0.f;
      for (int sg = 0; sg < nseg; ++sg) total += S.m_sum[sg];
      const float target = u01 * total;
      int tseg = -1, last_nonempty = -1;
      float toff = 0.f, last_off = 0.f, acc = 0.f;
      for (int sg = 0; sg < nseg; ++sg) {
        const float v = S.m_sum[sg];
        if (v > 0.f) { last_nonempty = sg; last_off = acc; }
        if (tseg < 0 && v > 0.f && target < acc + v) { tseg = sg; toff = acc; }
        acc += v;
      }
      if (tseg < 0) { tseg = last_nonempty; toff = last_off; }  // rounding: target fell past the end
      S.s_bin = tseg;
      S.s_before_mass = toff;
      S.s_tot_mass = target;
    }
    __syncthreads();
    const int tseg = S.s_bin;
    const float toff = S.s_before_mass, target = S.s_tot_mass;
    if (tseg == seg) {  // this warp owns the target segment
      float acc2 = toff;
      int found = -1, last_kept = -1;
      for (int base = seg_lo; base < seg_hi && found < 0; base += 128) {
        const int i0 = base + lane * 4;
        float x[4], w[4];
        if (vec && i0 + 3 < seg_hi) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(row + i0));
          x[0] = t4.x; x[1] = t4.y; x[2] = t4.z; x[3] = t4.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) x[j] = (i0 + j < seg_hi) ? row[i0 + j] : -INFINITY;
        }
        float lsum = 0.f;
        int lk = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float wj = 0.f;
          const bool kept = (x[j] != -INFINITY) && is_kept(x[j], wj);
          w[j] = kept ? wj : -1.f;  // -1 marks "not kept"
          if (kept) { lk = i0 + j; lsum += wj; }
        }
        float inc = lsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const bool hit = (lk >= 0) && (target < acc2 + inc);
        const uint32_t hits = __ballot_sync(0xffffffffu, hit);
        const uint32_t keeps = __ballot_sync(0xffffffffu, lk >= 0);
        if (keeps) last_kept = __shfl_sync(0xffffffffu, lk, 31 - __clz(keeps));
        if (hits) {
          const int hl = __ffs(hits) - 1;
          float a = acc2 + inc - lsum;  // mass before this lane's elements
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
        acc2 += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (found < 0) found = last_kept;
      if (lane == 0) out[row_idx] = found;
    }
  }
  cluster.sync();  // keep shared memory alive until every CTA has read the segment sums
  if (trc) g_topp_trace[5] = clock64();
}


// -------------------------------------------------------------------------------------------------------------------
// top-p by rejection (used when the caller does not ask for the kept-set size): a token drawn from the FULL softmax and
// accepted iff it lies in the kept set {t : mass of strictly more probable tokens <= top_p * Z} is distributed exactly as
// the renormalised kept set of inference.py:90-106.  The kept set carries >= top_p of the mass, so a candidate is accepted
// with probability >= top_p; four independent candidates are verified in one pass (all four rejected: <= 1e-4 at
// top_p = 0.9, then four more are drawn).  Three passes over the row, no histogram, no shared-memory atomics:
//   P0 row max   P1 per-warp segment masses (-> Z, inverse-CDF draw of the candidates, their warps walk 1/64 of the row)
//   P2 mass strictly above each candidate's logit -> first accepted candidate wins.
// One cluster of R CTAs per row; every element is weighted by ONE float expression in all passes.
// -------------------------------------------------------------------------------------------------------------------
struct RejShared {
  BlockRed red;
  float seg[32];        // this CTA's per-warp segment masses (read remotely)
  float all_seg[256];   // every segment of the row, gathered locally
  float cta_max;        // published
  float cand_x[4];      // candidate logits (written by the owning warp into EVERY rank)
  int cand_i[4];        // candidate token ids
  float mass[4];        // this CTA's mass strictly above each candidate (published)
  float target[4], toff[4];
  int tseg[4];
  float z;
  int accepted;
};

__global__ void __launch_bounds__(1024) sample_top_p_rej_kernel(const float* __restrict__ logits, long long ld, int* __restrict__ out,
                                                                int V, float inv_temp, float top_p, unsigned long long seed,
                                                                const int* __restrict__ step_ptr) {
  __shared__ RejShared S;
  cg::cluster_group cluster = cg::this_cluster();
  const int R = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const int row_idx = blockIdx.x / R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ row = logits + row_idx * ld;
  griddep_wait();
  if (threadIdx.x == 0) griddep_launch_dependents();

  const bool vec = ((V & 3) == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  const int nseg = R * 32;
  const int seg_len = ((V + nseg - 1) / nseg + 127) / 128 * 128;
  const int seg = rank * 32 + warp;
  const int seg_lo = min(V, seg * seg_len), seg_hi = min(V, seg_lo + seg_len);
  auto remote = [&](auto* ptr, int r) { return cluster.map_shared_rank(ptr, r); };

  // ---- P0: row max ----
  float mx = -INFINITY;
  for_each_in_segment(row, seg_lo, seg_hi, vec, lane, [&](float v, int) { mx = fmaxf(mx, v); });
  {
    const float t = block_reduce_max(mx, S.red);
    if (tid == 0) S.cta_max = t;
    cluster.sync();
    float acc = -INFINITY;
    for (int r = 0; r < R; ++r) acc = fmaxf(acc, *remote(&S.cta_max, r));
    mx = acc;
  }
  const float c = inv_temp * 1.4426950408889634f;  // w = exp((x - mx) * inv_temp) = exp2(x * c - mx * c)
  const float cm = mx * c;
  auto w_of = [&](float x) { return exp2f(fmaf(x, c, -cm)); };

  // ---- P1: segment masses ----
  {
    float m = 0.f;
    for_each_in_segment(row, seg_lo, seg_hi, vec, lane, [&](float x, int) { m += w_of(x); });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    if (lane == 0) S.seg[warp] = m;
  }
  cluster.sync();
  if (tid < nseg) S.all_seg[tid] = remote(S.seg, tid >> 5)[tid & 31];
  __syncthreads();
  if (tid == 0) {
    float z = 0.f;
    for (int sg = 0; sg < nseg; ++sg) z += S.all_seg[sg];
    S.z = z;
    S.accepted = -1;
  }
  __syncthreads();
  const float Z = S.z;
  const int step = step_ptr ? *step_ptr : 0;
  const uint64_t base_rnd = splitmix64(seed ^ splitmix64((static_cast<uint64_t>(step) << 32) | static_cast<uint32_t>(row_idx)));

  for (int round = 0; round < 16; ++round) {
    // ---- draw four candidates from the full distribution (inverse CDF over the segment masses, same in every rank) ----
    if (tid < 4) {
      const uint64_t rnd = splitmix64(base_rnd + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(4 * round + tid + 1));
      const float u01 = (static_cast<float>(rnd >> 40) + 0.5f) * (1.0f / 16777216.0f);
      const float target = u01 * Z;
      int tseg = -1, last_nonempty = -1;
      float toff = 0.f, last_off = 0.f, acc = 0.f;
      for (int sg = 0; sg < nseg; ++sg) {
        const float v = S.all_seg[sg];
        if (v > 0.f) { last_nonempty = sg; last_off = acc; }
        if (tseg < 0 && v > 0.f && target < acc + v) { tseg = sg; toff = acc; }
        acc += v;
      }
      if (tseg < 0) { tseg = last_nonempty; toff = last_off; }
      S.tseg[tid] = tseg;
      S.toff[tid] = toff;
      S.target[tid] = target;
    }
    __syncthreads();
    // ---- the warps that own the target segments walk them; the result goes to every rank ----
    for (int k = 0; k < 4; ++k) {
      if (S.tseg[k] != seg) continue;  // warp-uniform
      const float target = S.target[k];
      float acc2 = S.toff[k];
      int found = -1, last_el = -1;
      for (int base = seg_lo; base < seg_hi && found < 0; base += 128) {
        const int i0 = base + lane * 4;
        float w[4];
        float lsum = 0.f;
        int lk = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = i0 + j < seg_hi;
          w[j] = ok ? w_of(row[i0 + j]) : -1.f;  // -1 marks "no element"
          if (ok) { lk = i0 + j; lsum += w[j]; }
        }
        float inc = lsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const bool hit = (lk >= 0) && (target < acc2 + inc);
        const uint32_t hits = __ballot_sync(0xffffffffu, hit);
        const uint32_t has = __ballot_sync(0xffffffffu, lk >= 0);
        if (has) last_el = __shfl_sync(0xffffffffu, lk, 31 - __clz(has));
        if (hits) {
          const int hl = __ffs(hits) - 1;
          float a = acc2 + inc - lsum;
          int f = -1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (f < 0 && w[j] >= 0.f) {
              a += w[j];
              if (target < a) f = i0 + j;
            }
          }
          if (f < 0) f = lk;
          found = __shfl_sync(0xffffffffu, f, hl);
        }
        acc2 += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (found < 0) found = last_el;
      if (lane < R) {
        *remote(&S.cand_x[k], lane) = row[found];
        *remote(&S.cand_i[k], lane) = found;
      }
    }
    cluster.sync();
    // ---- P2: mass strictly above each candidate ----
    const float cx0 = S.cand_x[0], cx1 = S.cand_x[1], cx2 = S.cand_x[2], cx3 = S.cand_x[3];
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
    for_each_in_segment(row, seg_lo, seg_hi, vec, lane, [&](float x, int) {
      const float w = w_of(x);
      m0 += x > cx0 ? w : 0.f;
      m1 += x > cx1 ? w : 0.f;
      m2 += x > cx2 ? w : 0.f;
      m3 += x > cx3 ? w : 0.f;
    });
    {
      const float t0 = block_reduce_sum(m0, S.red), t1 = block_reduce_sum(m1, S.red);
      const float t2 = block_reduce_sum(m2, S.red), t3 = block_reduce_sum(m3, S.red);
      if (tid == 0) { S.mass[0] = t0; S.mass[1] = t1; S.mass[2] = t2; S.mass[3] = t3; }
    }
    cluster.sync();
    if (tid == 0) {
      int acc_k = -1;
      for (int k = 0; k < 4 && acc_k < 0; ++k) {
        float above = 0.f;
        for (int r = 0; r < R; ++r) above += remote(S.mass, r)[k];
        if (above <= top_p * Z) acc_k = k;  // (the most probable token always passes: nothing lies above it)
      }
      S.accepted = acc_k;
    }
    __syncthreads();
    const int acc_k = S.accepted;
    cluster.sync();  // every rank has read the published masses / candidates before the next round overwrites them
    if (acc_k >= 0) {
      if (tid == 0 && rank == 0) out[row_idx] = S.cand_i[acc_k];
      return;
    }
  }
  // ---- every candidate of every round rejected (tiny top_p on a flat row): emit the most probable token, which is always
  //      in the kept set (nothing lies above it); lowest index among exact ties ----
  int first = 0x7fffffff;
  for_each_in_segment(row, seg_lo, seg_hi, vec, lane, [&](float x, int i) { if (x == mx) first = min(first, i); });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
  if (tid == 0) S.accepted = 0x7fffffff;
  __syncthreads();
  if (lane == 0) atomicMin(&S.accepted, first);
  cluster.sync();
  if (tid == 0 && rank == 0) {
    int best = 0x7fffffff;
    for (int r = 0; r < R; ++r) best = min(best, *remote(&S.accepted, r));
    out[row_idx] = best == 0x7fffffff ? 0 : best;
  }
  cluster.sync();  // no rank leaves while its shared memory may still be read
}


// -------------------------------------------------------------------------------------------------------------------
// Samplers fed by the lm_head epilogue's segment statistics (gemm_tcgen05.cu: swap_tile_epilogue_f32_stats): for every
// 32-token vocabulary segment g of a row, stats[g] = (m_g, s_g) = (max logit, sum exp2((x - m_g) * c)), c = inv_temp * log2 e.
// Row maximum M = max m_g, segment mass w_g = s_g * exp2((m_g - M) * c), partition function Z = sum w_g: the first two of
// the three passes of the rejection sampler above now read 64 KB of statistics instead of the 1 MB logit row; only the
// verification pass (mass of strictly more probable tokens, inference.py:96-100) still walks the row.
// -------------------------------------------------------------------------------------------------------------------
constexpr int kStatMaxPer = 8;  // segments per thread (host: nseg <= kStatMaxPer * R * 1024)

struct StatShared {
  BlockRed red;
  float wmass[32];      // this CTA's per-warp masses (read remotely)
  float all_w[256];     // every warp of the row (cluster), vocabulary order
  float cta_max;        // published
  float cand_x[4];      // candidate logits (written by the owning warp into EVERY rank)
  int cand_i[4];
  float mass[4];        // this CTA's mass strictly above each candidate (published)
  float target[4], toff[4];
  int twarp[4];
  float z;
  int accepted;
  int max_seg;          // lowest segment holding the row maximum (fallback: the most probable token is always kept)
  float red4[32][4];
};

__global__ void __launch_bounds__(1024) sample_top_p_stats_kernel(const float* __restrict__ logits, long long ld,
                                                                  const float2* __restrict__ stats, long long stats_ld, int nseg,
                                                                  int* __restrict__ out, int V, float inv_temp, float top_p,
                                                                  unsigned long long seed, const unsigned long long* __restrict__ seed_ptr,
                                                                  const int* __restrict__ step_ptr) {
  __shared__ StatShared S;
  cg::cluster_group cluster = cg::this_cluster();
  const int R = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const int row_idx = blockIdx.x / R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ row = logits + row_idx * ld;
  const float2* __restrict__ st = stats + row_idx;  // segment g of this row: st[g * stats_ld]
  griddep_wait();
  if (threadIdx.x == 0) griddep_launch_dependents();
  auto remote = [&](auto* ptr, int r) { return cluster.map_shared_rank(ptr, r); };
  const float c = inv_temp * 1.4426950408889634f;

  // ---- this thread's contiguous run of segments (threads in cluster order = vocabulary order) ----
  const int gt = rank * 1024 + tid;
  const int per = (nseg + R * 1024 - 1) / (R * 1024);
  const int s_lo = min(nseg, gt * per), s_hi = min(nseg, s_lo + per);
  float2 my[kStatMaxPer];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < kStatMaxPer; ++k) {
    my[k] = (k < per && s_lo + k < s_hi) ? __ldcg(st + (s_lo + k) * stats_ld) : make_float2(-INFINITY, 0.f);
    if (my[k].y > 0.f) mx = fmaxf(mx, my[k].x);
  }
  if (tid == 0) { S.accepted = -1; S.max_seg = 0x7fffffff; }
  {
    const float t = block_reduce_max(mx, S.red);
    if (tid == 0) S.cta_max = t;
    cluster.sync();
    float acc = -INFINITY;
    for (int r = 0; r < R; ++r) acc = fmaxf(acc, *remote(&S.cta_max, r));
    mx = acc;
  }
  const float cm = mx * c;
  auto w_of = [&](float x) { return exp2f(fmaf(x, c, -cm)); };  // ONE expression for every element-level weight
  float wk[kStatMaxPer];
  float tm = 0.f;
#pragma unroll
  for (int k = 0; k < kStatMaxPer; ++k) {
    wk[k] = my[k].y > 0.f ? my[k].y * exp2f((my[k].x - mx) * c) : 0.f;
    tm += wk[k];
    if (my[k].y > 0.f && my[k].x == mx) atomicMin(&S.max_seg, s_lo + k);
  }
  {
    float m = tm;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    if (lane == 0) S.wmass[warp] = m;
  }
  cluster.sync();
  const int nwarp = R * 32;
  if (tid < nwarp) S.all_w[tid] = remote(S.wmass, tid >> 5)[tid & 31];
  __syncthreads();
  if (tid == 0) {
    float z = 0.f;
    for (int w = 0; w < nwarp; ++w) z += S.all_w[w];
    S.z = z;
  }
  __syncthreads();
  const float Z = S.z;
  const int step = step_ptr ? *step_ptr : 0;
  const unsigned long long sd = seed_ptr ? *seed_ptr : seed;
  const uint64_t base_rnd = splitmix64(sd ^ splitmix64((static_cast<uint64_t>(step) << 32) | static_cast<uint32_t>(row_idx)));

  // the verification pass walks the row in per-warp element ranges (same split as the rejection kernel above)
  const bool vec = ((V & 3) == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  const int seg_len = ((V + nwarp - 1) / nwarp + 127) / 128 * 128;
  const int e_lo = min(V, (rank * 32 + warp) * seg_len), e_hi = min(V, e_lo + seg_len);

  int acc_k = -1;
  for (int round = 0; round < 16 && acc_k < 0; ++round) {
    // ---- four candidates from the full distribution: warp by inverse CDF over the warp masses (same in every rank) ----
    if (tid < 4) {
      const uint64_t rnd = splitmix64(base_rnd + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(4 * round + tid + 1));
      const float u01 = (static_cast<float>(rnd >> 40) + 0.5f) * (1.0f / 16777216.0f);
      const float target = u01 * Z;
      int tw = -1, last_nonempty = -1;
      float toff = 0.f, last_off = 0.f, acc = 0.f;
      for (int w = 0; w < nwarp; ++w) {
        const float v = S.all_w[w];
        if (v > 0.f) { last_nonempty = w; last_off = acc; }
        if (tw < 0 && v > 0.f && target < acc + v) { tw = w; toff = acc; }
        acc += v;
      }
      if (tw < 0) { tw = last_nonempty; toff = last_off; }
      S.twarp[tid] = tw;
      S.toff[tid] = toff;
      S.target[tid] = target;
    }
    __syncthreads();
    // ---- the owning warp: lane by prefix of the thread masses, segment inside the lane's run, token inside the segment ----
    for (int k = 0; k < 4; ++k) {
      if (S.twarp[k] != rank * 32 + warp) continue;  // warp-uniform
      const float target = S.target[k];
      float inc = tm;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const float before = S.toff[k] + inc - tm;
      const uint32_t hits = __ballot_sync(0xffffffffu, tm > 0.f && target < before + tm);
      const uint32_t has = __ballot_sync(0xffffffffu, tm > 0.f);
      const int hl = hits ? __ffs(hits) - 1 : (has ? 31 - __clz(has) : 0);
      int sseg = s_lo;
      float soff = before;
      if (lane == hl) {
        float a = before;
        int pick = -1, lastk = 0;
        float lastoff = before;
#pragma unroll
        for (int kk = 0; kk < kStatMaxPer; ++kk) {
          if (wk[kk] > 0.f) {
            lastk = kk; lastoff = a;
            if (pick < 0 && target < a + wk[kk]) { pick = kk; soff = a; }
            a += wk[kk];
          }
        }
        if (pick < 0) { pick = lastk; soff = lastoff; }
        sseg = s_lo + pick;
      }
      sseg = __shfl_sync(0xffffffffu, sseg, hl);
      soff = __shfl_sync(0xffffffffu, soff, hl);
      const int idx = sseg * 32 + lane;
      const bool ok = idx < V;
      const float x = ok ? __ldcg(row + idx) : -INFINITY;
      const float w = ok ? w_of(x) : 0.f;
      float winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      const uint32_t ehits = __ballot_sync(0xffffffffu, ok && w > 0.f && target < soff + winc);
      const uint32_t ehas = __ballot_sync(0xffffffffu, ok && w > 0.f);
      const uint32_t eany = __ballot_sync(0xffffffffu, ok);
      const int el = ehits ? __ffs(ehits) - 1 : (ehas ? 31 - __clz(ehas) : (eany ? 31 - __clz(eany) : 0));
      const float fx = __shfl_sync(0xffffffffu, x, el);
      if (lane < R) {
        *remote(&S.cand_x[k], lane) = fx;
        *remote(&S.cand_i[k], lane) = min(V - 1, sseg * 32 + el);
      }
    }
    cluster.sync();
    // ---- mass strictly above each candidate ----
    const float cx0 = S.cand_x[0], cx1 = S.cand_x[1], cx2 = S.cand_x[2], cx3 = S.cand_x[3];
    float mm[4] = {0.f, 0.f, 0.f, 0.f};
    for_each_in_segment(row, e_lo, e_hi, vec, lane, [&](float x, int) {
      const float w = w_of(x);
      mm[0] += x > cx0 ? w : 0.f;
      mm[1] += x > cx1 ? w : 0.f;
      mm[2] += x > cx2 ? w : 0.f;
      mm[3] += x > cx3 ? w : 0.f;
    });
    {
      float tot[4];
      block_reduce_sum4(mm, S.red4, tot);
      if (tid == 0) { S.mass[0] = tot[0]; S.mass[1] = tot[1]; S.mass[2] = tot[2]; S.mass[3] = tot[3]; }
    }
    cluster.sync();
    if (tid == 0) {
      int a = -1;
      for (int k = 0; k < 4 && a < 0; ++k) {
        float above = 0.f;
        for (int r = 0; r < R; ++r) above += remote(S.mass, r)[k];
        if (above <= top_p * Z) a = k;  // (the most probable token always passes: nothing lies above it)
      }
      S.accepted = a;
    }
    __syncthreads();
    acc_k = S.accepted;
    cluster.sync();  // every rank has read the published masses / candidates before the next round overwrites them
  }
  if (acc_k >= 0) {
    if (tid == 0 && rank == 0) out[row_idx] = S.cand_i[acc_k];
    return;
  }
  // ---- every candidate of every round rejected (tiny top_p on a flat row): the most probable token, which is always in
  //      the kept set; lowest index among exact ties ----
  int ms = 0x7fffffff;
  for (int r = 0; r < R; ++r) ms = min(ms, *remote(&S.max_seg, r));
  cluster.sync();  // no rank leaves while its shared memory may still be read
  if (rank == 0 && warp == 0) {
    const int idx = ms * 32 + lane;
    const bool hit = ms != 0x7fffffff && idx < V && __ldcg(row + idx) == mx;
    const uint32_t b = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) out[row_idx] = b ? ms * 32 + __ffs(b) - 1 : 0;
  }
}

// Greedy argmax (inference.py:67-68) from the segment statistics: lowest segment holding the maximum, then the first of its
// 32 logits equal to it (ties -> lowest index, as torch.argmax over the row).
__global__ void __launch_bounds__(256) argmax_stats_kernel(const float* __restrict__ logits, long long ld,
                                                           const float2* __restrict__ stats, long long stats_ld, int nseg,
                                                           int* __restrict__ out, int V) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const float* __restrict__ row = logits + blockIdx.x * ld;
  const float2* __restrict__ st = stats + blockIdx.x;  // segment g of this row: st[g * stats_ld]
  griddep_wait();
  if (threadIdx.x == 0) griddep_launch_dependents();
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int g = threadIdx.x; g < nseg; g += 256) {
    const float2 v = __ldcg(st + g * stats_ld);
    if (v.y > 0.f && (v.x > best || (v.x == best && g < bi))) { best = v.x; bi = g; }
  }
  auto merge = [&](float ov, int oi) { if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; } };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) merge(__shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < 8 ? sv[lane] : -INFINITY;
    bi = lane < 8 ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) merge(__shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o));
    best = __shfl_sync(0xffffffffu, best, 0);
    bi = __shfl_sync(0xffffffffu, bi, 0);
    const int idx = bi * 32 + lane;
    const bool hit = bi != 0x7fffffff && idx < V && __ldcg(row + idx) == best;
    const uint32_t b = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) out[blockIdx.x] = b ? bi * 32 + __ffs(b) - 1 : 0;
  }
}

}  // namespace pg

using namespace pg;

extern "C" int pg_debug_topp_trace(long long* host_out16) {
  return cudaMemcpyFromSymbol(host_out16, pg::g_topp_trace, 16 * sizeof(long long)) == cudaSuccess ? 0 : -2;
}

extern "C" int pg_debug_set_topp_bracket(int d) {
  return cudaMemcpyToSymbol(pg::g_topp_bracket, &d, sizeof(int)) == cudaSuccess ? 0 : -2;
}

extern "C" int pg_debug_topp_retries(void) {
  int v = 0;
  cudaMemcpyFromSymbol(&v, pg::g_topp_retries, sizeof(int));
  return v;
}

extern "C" int pg_argmax(const float* logits, long long ld, int* out, int B, int V, void* stream) {
  if (B <= 0 || V <= 0) return PG_ERR_ARG;
  return launch_kernel(argmax_kernel, dim3(B), dim3(1024), 0, reinterpret_cast<cudaStream_t>(stream), logits, ld, out, V) == cudaSuccess
             ? PG_OK : PG_ERR_CUDA;
}

extern "C" int pg_sample_top_p(const float* logits, long long ld, int* out, int* kept_count, int B, int V,
                               float inv_temperature, float top_p, unsigned long long seed, const int* step_ptr,
                               void* stream) {
  if (B <= 0 || V <= 0 || !(inv_temperature > 0.f) || !(top_p >= 0.f)) return PG_ERR_ARG;
  // cluster of R CTAs per row: enough CTAs to cover the GPU when the batch alone cannot
  int R = 1;
  if (V >= 65536) R = 2;  // (64 regs x 1024 threads = one CTA per SM: 2 x 64 rows covers 128 of the 148 SMs)
  const bool rejection = kept_count == nullptr && top_p > 0.f;  // the histogram-select kernel reports the kept-set size
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(B) * R);
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = R;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pg_pdl_enabled() ? 2 : 1;
  pg_count_launch(1);
  if (rejection)
    return cudaLaunchKernelEx(&cfg, sample_top_p_rej_kernel, logits, ld, out, V, inv_temperature, top_p, seed, step_ptr) == cudaSuccess
               ? PG_OK : PG_ERR_CUDA;
  return cudaLaunchKernelEx(&cfg, sample_top_p_kernel, logits, ld, out, kept_count, V, inv_temperature, top_p, seed, step_ptr) ==
                 cudaSuccess
             ? PG_OK
             : PG_ERR_CUDA;
}

static int g_stats_cluster = 0;  // tuning sweeps only: forces the cluster size of pg_sample_top_p_stats (0 = automatic)
extern "C" int pg_debug_set_sampler_cluster(int r) {
  if (r != 0 && r != 1 && r != 2 && r != 4 && r != 8) return PG_ERR_ARG;
  g_stats_cluster = r;
  return 0;
}

extern "C" int pg_argmax_stats(const float* logits, long long ld, const void* stats, long long stats_ld, int* out, int B, int V,
                               void* stream) {
  if (B <= 0 || V <= 0 || stats == nullptr || stats_ld < B) return PG_ERR_ARG;
  return launch_kernel(argmax_stats_kernel, dim3(B), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), logits, ld,
                       static_cast<const float2*>(stats), stats_ld, (V + 31) / 32, out, V) == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}

extern "C" int pg_sample_top_p_stats(const float* logits, long long ld, const void* stats, long long stats_ld, int* out, int B,
                                     int V, float inv_temperature, float top_p, unsigned long long seed,
                                     const unsigned long long* seed_ptr, const int* step_ptr, void* stream) {
  if (B <= 0 || V <= 0 || !(inv_temperature > 0.f) || !(top_p >= 0.f) || stats == nullptr) return PG_ERR_ARG;
  const int nseg = (V + 31) / 32;
  if (stats_ld < B) return PG_ERR_ARG;
  // CTAs per row: the verification pass is split over the cluster, every split costs cluster barriers (~1.5 us each, four
  // per round); fill the GPU once when the batch alone does not
  int R = 1;
  if (g_stats_cluster > 0) R = g_stats_cluster;
  else if (V >= 65536) { while (R < 8 && 2 * R * B <= pg::num_sms()) R *= 2; }
  while (nseg > kStatMaxPer * R * 1024 && R < 8) R *= 2;
  if (nseg > kStatMaxPer * R * 1024) return PG_ERR_ARG;  // > 2 M-token vocabularies: use pg_sample_top_p
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(B) * R);
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = R;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pg_pdl_enabled() ? 2 : 1;
  pg_count_launch(1);
  return cudaLaunchKernelEx(&cfg, sample_top_p_stats_kernel, logits, ld, static_cast<const float2*>(stats), stats_ld, nseg, out, V,
                            inv_temperature, top_p, seed, seed_ptr, step_ptr) == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}
