// HBM-bound row kernels around the GEMMs: LayerNorm, GemmaRMSNorm, im2col, embedding merge,
// RoPE + paged KV append, KV gather, weight packing.  All are single-pass, 16-byte vectorised where the layout allows.
#include "common.cuh"
#include "paligemma_b200.h"

namespace pg {

typedef __nv_bfloat16 bf16;

PG_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result broadcast to every thread; `red` holds >= 33 floats
PG_DEVINL float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// ---------------------------------------------------------------------------------------------------------
// nn.LayerNorm (modeling_siglip.py:199-204,310,319): biased variance, eps inside the sqrt, affine
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, bf16* __restrict__ y_bf,
                                                        float* __restrict__ y_f, int D, float eps) {
  extern __shared__ float row[];  // D floats + 33
  float* red = row + D;
  const long long r = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x + r * D);
  float s = 0.f;
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    float4 v = xr[i];
    reinterpret_cast<float4*>(row)[i] = v;
    s += v.x + v.y + v.z + v.w;
  }
  const float mean = block_sum(s, red) / D;
  float ss = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float d = row[i] - mean;
    ss += d * d;
  }
  const float rstd = rsqrtf(block_sum(ss, red) / D + eps);
  for (int i = threadIdx.x; i < D / 2; i += blockDim.x) {
    float a = (row[2 * i] - mean) * rstd * gamma[2 * i] + beta[2 * i];
    float b = (row[2 * i + 1] - mean) * rstd * gamma[2 * i + 1] + beta[2 * i + 1];
    if (y_bf) reinterpret_cast<uint32_t*>(y_bf + r * D)[i] = pack_bf16(a, b);
    if (y_f) reinterpret_cast<float2*>(y_f + r * D)[i] = make_float2(a, b);
  }
}

// ---------------------------------------------------------------------------------------------------------
// GemmaRMSNorm (modeling_gemma.py:172-181): fp32, x * rsqrt(mean(x^2) + eps) * (1 + w)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      bf16* __restrict__ y, int D, float eps, int early_trigger) {
  extern __shared__ float row[];
  float* red = row + D;
  const long long r = blockIdx.x;
  // early_trigger (decode chain): trigger BEFORE the wait, so that the GEMM consuming this norm becomes resident -- barriers,
  // TMEM, tensor maps, first weight stages -- while the producer of x is still running, one kernel earlier than with the usual
  // wait-then-trigger order.  Safe: the consumer's own griddepcontrol.wait waits for THIS kernel to complete, which has
  // waited for its producer; before its wait the consumer only touches weights.
  if (early_trigger && threadIdx.x == 0) griddep_launch_dependents();
  griddep_wait();
  if (!early_trigger && threadIdx.x == 0) griddep_launch_dependents();
  const float4* xr = reinterpret_cast<const float4*>(x + r * D);
  float ss = 0.f;
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    float4 v = xr[i];
    reinterpret_cast<float4*>(row)[i] = v;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  const float rstd = rsqrtf(block_sum(ss, red) / D + eps);
  for (int i = threadIdx.x; i < D / 2; i += blockDim.x) {
    float a = row[2 * i] * rstd * (1.0f + w[2 * i]);
    float b = row[2 * i + 1] * rstd * (1.0f + w[2 * i + 1]);
    reinterpret_cast<uint32_t*>(y + r * D)[i] = pack_bf16(a, b);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Warp-per-row variants for the prefill (tens of thousands of short rows): the row lives in registers (D/128 float4 per
// lane), the two reductions are warp shuffles, nothing goes through shared memory or a block barrier.  D % 128 == 0,
// D <= 128 * NV * ... (NV float4 per lane).
// ---------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) layernorm_warp_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, bf16* __restrict__ y_bf,
                                                             float* __restrict__ y_f, int rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int nv = D >> 7;  // float4 per lane
  const float4* xr = reinterpret_cast<const float4*>(x + r * D);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nv) {
      v[i] = xr[lane + 32 * i];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(s) / D;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += a * a + b * b + c * c + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / D + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nv) {
      const int c4 = lane + 32 * i;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), bt = __ldg(reinterpret_cast<const float4*>(beta) + c4);
      const float a = (v[i].x - mean) * rstd * g.x + bt.x, b = (v[i].y - mean) * rstd * g.y + bt.y;
      const float c = (v[i].z - mean) * rstd * g.z + bt.z, d = (v[i].w - mean) * rstd * g.w + bt.w;
      if (y_bf) reinterpret_cast<uint2*>(y_bf + r * D)[c4] = make_uint2(pack_bf16(a, b), pack_bf16(c, d));
      if (y_f) reinterpret_cast<float4*>(y_f + r * D)[c4] = make_float4(a, b, c, d);
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256) rmsnorm_warp_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           bf16* __restrict__ y, int rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int nv = D >> 7;
  const float4* xr = reinterpret_cast<const float4*>(x + r * D);
  float4 v[NV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nv) {
      v[i] = xr[lane + 32 * i];
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / D + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nv) {
      const int c4 = lane + 32 * i;
      const float4 g = __ldg(reinterpret_cast<const float4*>(w) + c4);
      reinterpret_cast<uint2*>(y + r * D)[c4] = make_uint2(pack_bf16(v[i].x * rstd * (1.0f + g.x), v[i].y * rstd * (1.0f + g.y)),
                                                          pack_bf16(v[i].z * rstd * (1.0f + g.z), v[i].w * rstd * (1.0f + g.w)));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// im2col for the patch-embedding conv (modeling_siglip.py:258-263): column = c*P*P + ky*P + kx
// ---------------------------------------------------------------------------------------------------------
__global__ void im2col_kernel(const float* __restrict__ px, bf16* __restrict__ out, int C, int H, int W, int P, int Kpad,
                              long long total) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int col = static_cast<int>(idx % Kpad);
  const long long prow = idx / Kpad;
  const int Wp = W / P, Hp = H / P;
  const int n = static_cast<int>(prow % (Hp * Wp));
  const long long b = prow / (Hp * Wp);
  float v = 0.f;
  if (col < C * P * P) {
    const int c = col / (P * P), ky = (col / P) % P, kx = col % P;
    const int yy = (n / Wp) * P + ky, xx = (n % Wp) * P + kx;
    v = px[((b * C + c) * H + yy) * W + xx];
  }
  out[idx] = __float2bfloat16(v);
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 2;
  if (i + 1 < n) {
    const float2 v = *reinterpret_cast<const float2*>(src + i);
    *reinterpret_cast<uint32_t*>(dst + i) = pack_bf16(v.x, v.y);
  } else if (i < n) {
    dst[i] = __float2bfloat16(src[i]);
  }
}

// packed row p: block = p / 128, within = p % 128; within < 64 -> gate[block*64 + within] else up[block*64 + within-64]
__global__ void pack_gate_up_kernel(const bf16* __restrict__ gate, const bf16* __restrict__ up, bf16* __restrict__ packed,
                                    int F, int K8) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = 2ll * F * K8;
  if (idx >= total) return;
  const long long p = idx / K8;
  const int c = static_cast<int>(idx % K8);
  const long long blk = p / 128;
  const int within = static_cast<int>(p % 128);
  const bf16* src = within < 64 ? gate + (blk * 64 + within) * K8 * 8 : up + (blk * 64 + within - 64) * K8 * 8;
  reinterpret_cast<uint4*>(packed)[idx] = reinterpret_cast<const uint4*>(src)[c];
}

// ---------------------------------------------------------------------------------------------------------
// embedding merge (modeling_paligemma.py:93-128,195,288; modeling_gemma.py:510-511)
// ---------------------------------------------------------------------------------------------------------
// pass 1: one block per batch row; src[b,s] = j >= 0 (j-th image token) | -1 text | -2 pad; pos = cumsum(mask), 1 at mask==0
__global__ void __launch_bounds__(1024) merge_scan_kernel(const long long* __restrict__ ids, const long long* __restrict__ mask,
                                                          int* __restrict__ src, int* __restrict__ pos, int* __restrict__ err,
                                                          int* __restrict__ dst_row, int S, int N, long long image_token,
                                                          long long pad_token) {
  __shared__ int wsum_img[32], wsum_msk[32];
  __shared__ int carry_img, carry_msk;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { carry_img = 0; carry_msk = 0; }
  __syncthreads();
  for (int base = 0; base < S; base += blockDim.x) {
    const int s = base + threadIdx.x;
    long long id = 0, mk = 0;
    if (s < S) { id = ids[static_cast<long long>(b) * S + s]; mk = mask[static_cast<long long>(b) * S + s]; }
    const int is_img = (s < S && id == image_token) ? 1 : 0;
    const int mv = (s < S) ? static_cast<int>(mk) : 0;
    int xi = is_img, xm = mv;  // inclusive warp scans
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int ti = __shfl_up_sync(0xffffffffu, xi, o), tm = __shfl_up_sync(0xffffffffu, xm, o);
      if (lane >= o) { xi += ti; xm += tm; }
    }
    if (lane == 31) { wsum_img[warp] = xi; wsum_msk[warp] = xm; }
    __syncthreads();
    int oi = carry_img, om = carry_msk;
    for (int w = 0; w < warp; ++w) { oi += wsum_img[w]; om += wsum_msk[w]; }
    if (s < S) {
      const int j = oi + xi - is_img;  // exclusive count of image tokens
      src[static_cast<long long>(b) * S + s] = (id == pad_token) ? -2 : (is_img ? j : -1);
      // inverse map for the projector GEMM that scatters its rows: image feature j of row b lands on merged row b*S + s
      if (dst_row != nullptr && is_img && id != pad_token && j < N) dst_row[static_cast<long long>(b) * N + j] = b * S + s;
      pos[static_cast<long long>(b) * S + s] = (mk == 0) ? 1 : (om + xm);
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) { carry_img = oi + xi; carry_msk = om + xm; }
    __syncthreads();
  }
  if (threadIdx.x == 0 && carry_img != N) atomicExch(err, 1);
}

// pass 2: one block per token
__global__ void __launch_bounds__(256) merge_gather_kernel(const long long* __restrict__ ids, const int* __restrict__ src,
                                                           const bf16* __restrict__ embed, const float* __restrict__ img,
                                                           float* __restrict__ h, int S, int D, int N, float text_scale,
                                                           float img_scale, int skip_image) {
  const long long tok = blockIdx.x;
  const int b = static_cast<int>(tok / S);
  const int sidx = src[tok];
  float4* dst = reinterpret_cast<float4*>(h + tok * D);
  if (sidx >= 0 && skip_image) return;  // the projector GEMM has written this row (pg_gemm_bf16_fused out_row_map)
  if (sidx == -2) {
    for (int i = threadIdx.x; i < D / 4; i += blockDim.x) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (sidx >= 0) {
    const int j = sidx < N ? sidx : N - 1;  // (err flag already raised when counts mismatch)
    const float4* s4 = reinterpret_cast<const float4*>(img + (static_cast<long long>(b) * N + j) * D);
    for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
      float4 v = s4[i];
      dst[i] = make_float4(v.x * img_scale, v.y * img_scale, v.z * img_scale, v.w * img_scale);
    }
  } else {
    const uint2* e = reinterpret_cast<const uint2*>(embed + ids[tok] * D);
    for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
      const uint2 u = e[i];
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      dst[i] = make_float4(a.x * text_scale, a.y * text_scale, c.x * text_scale, c.y * text_scale);
    }
  }
}

// decode-step embedding of the sampled token (same merge rules, q_len = 1)
__global__ void __launch_bounds__(256) embed_tokens_kernel(const int* __restrict__ tokens, const bf16* __restrict__ embed,
                                                           const float* __restrict__ img, float* __restrict__ h, int D,
                                                           int N, float text_scale, float img_scale, long long pad_token,
                                                           long long image_token) {
  const int b = blockIdx.x;
  griddep_wait();
  if (threadIdx.x == 0) griddep_launch_dependents();
  const long long id = tokens[b];
  float4* dst = reinterpret_cast<float4*>(h + static_cast<long long>(b) * D);
  if (id == pad_token) {
    for (int i = threadIdx.x; i < D / 4; i += blockDim.x) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (id == image_token && img != nullptr) {
    // masked_scatter with a single True position per row consumes the first feature row of that image
    const float4* s4 = reinterpret_cast<const float4*>(img + static_cast<long long>(b) * N * D);
    for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
      float4 v = s4[i];
      dst[i] = make_float4(v.x * img_scale, v.y * img_scale, v.z * img_scale, v.w * img_scale);
    }
  } else {
    const uint2* e = reinterpret_cast<const uint2*>(embed + id * D);
    for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
      const uint2 u = e[i];
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      dst[i] = make_float4(a.x * text_scale, a.y * text_scale, c.x * text_scale, c.y * text_scale);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// RoPE (rotate-half, modeling_gemma.py:116-151) + KVCache.update (:18-57) into pages
// ---------------------------------------------------------------------------------------------------------
template <bool F32IN>
__global__ void __launch_bounds__(128) rope_kv_append_kernel(const void* __restrict__ qkv_, const int* __restrict__ pos,
                                                             bf16* __restrict__ q_out, bf16* __restrict__ k_out,
                                                             bf16* __restrict__ v_out, bf16* __restrict__ k_pages,
                                                             bf16* __restrict__ v_pages, const int* __restrict__ page_table,
                                                             const int* __restrict__ slot_base, int tokens_per_seq, int Hq,
                                                             int Hkv, int dh, int page_size, int max_pages,
                                                             const float* __restrict__ inv_freq) {
  const long long t = blockIdx.x;
  const int b = static_cast<int>(t / tokens_per_seq);
  const int half = dh / 2;
  const int W = (Hq + 2 * Hkv) * dh;
  const float p = static_cast<float>(pos[t]);
  auto ld = [&](long long i) -> float {
    if (F32IN) return static_cast<const float*>(qkv_)[t * W + i];
    return __bfloat162float(static_cast<const bf16*>(qkv_)[t * W + i]);
  };
  long long kv_row = -1;
  if (k_pages != nullptr) {
    const int slot = slot_base[b] + static_cast<int>(t - static_cast<long long>(b) * tokens_per_seq);
    const int page = page_table[b * max_pages + slot / page_size];
    kv_row = (static_cast<long long>(page) * page_size + slot % page_size) * Hkv * dh;
  }
  // q and k heads: rotate pairs (i, i + dh/2); the angle depends on (token, i) only, so one sincos serves every head
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    float sn, cs;
    sincosf(p * inv_freq[i], &sn, &cs);
    for (int head = 0; head < Hq + Hkv; ++head) {
      const float x1 = ld(head * dh + i), x2 = ld(head * dh + i + half);
      const bf16 y1 = __float2bfloat16(x1 * cs - x2 * sn);
      const bf16 y2 = __float2bfloat16(x2 * cs + x1 * sn);
      if (head < Hq) {
        q_out[t * Hq * dh + head * dh + i] = y1;
        q_out[t * Hq * dh + head * dh + i + half] = y2;
      } else {
        const int hk = head - Hq;
        if (k_out) { k_out[t * Hkv * dh + hk * dh + i] = y1; k_out[t * Hkv * dh + hk * dh + i + half] = y2; }
        if (kv_row >= 0) { k_pages[kv_row + hk * dh + i] = y1; k_pages[kv_row + hk * dh + i + half] = y2; }
      }
    }
  }
  for (int idx = threadIdx.x; idx < Hkv * dh; idx += blockDim.x) {
    const bf16 v = __float2bfloat16(ld((Hq + Hkv) * dh + idx));
    if (v_out) v_out[t * Hkv * dh + idx] = v;
    if (kv_row >= 0) v_pages[kv_row + idx] = v;
  }
}

__global__ void kv_gather_kernel(const bf16* __restrict__ pages, const int* __restrict__ page_table, bf16* __restrict__ dense,
                                 int len, int Hkv, int dh, int page_size, int max_pages, long long total) {
  // dense layout [B, Hkv, len, dh] (the reference KVCache layout, modeling_gemma.py:285-287)
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int d = static_cast<int>(idx % dh);
  long long r = idx / dh;
  const int s = static_cast<int>(r % len); r /= len;
  const int hk = static_cast<int>(r % Hkv);
  const int b = static_cast<int>(r / Hkv);
  const int page = page_table[b * max_pages + s / page_size];
  dense[idx] = pages[((static_cast<long long>(page) * page_size + s % page_size) * Hkv + hk) * dh + d];
}

// counters is int32 [n_counters][B] (position ids, write slots, kv lengths): all advance by one token
__global__ void advance_decode_kernel(const int* __restrict__ next, int* __restrict__ tok_hist, int* __restrict__ cur_tok,
                                      int* __restrict__ counters, int n_counters, int* __restrict__ step, int B) {
  const int b = threadIdx.x;
  griddep_wait();
  if (threadIdx.x == 0) griddep_launch_dependents();
  const int st = *step;
  if (b < B) {
    if (tok_hist) tok_hist[static_cast<long long>(st) * B + b] = next[b];
    if (cur_tok) cur_tok[b] = next[b];
    for (int c = 0; c < n_counters; ++c) counters[c * B + b] += 1;
  }
  __syncthreads();  // single block: every thread has read *step
  if (threadIdx.x == 0) *step = st + 1;
}

// Continuous batching (slot = one row of the decode batch, refilled by the host between graph replays): the token log is a
// ring of `ring` steps, and a slot only advances while its kv length is below its budget kv_limit[b] (prompt + new tokens).
// A slot at its budget (finished or idle) is frozen: it keeps recomputing its last position, which is harmless.
__global__ void advance_decode_slots_kernel(const int* __restrict__ next, int* __restrict__ tok_ring, int ring,
                                            int* __restrict__ cur_tok, int* __restrict__ counters,
                                            const int* __restrict__ kv_limit, int* __restrict__ step, int B) {
  const int b = threadIdx.x;
  griddep_wait();
  if (threadIdx.x == 0) griddep_launch_dependents();
  const int st = *step;
  if (b < B) {
    const int t = next[b];
    tok_ring[static_cast<long long>(st % ring) * B + b] = t;
    cur_tok[b] = t;
    if (counters[2 * B + b] < kv_limit[b]) {
      counters[b] += 1;
      counters[B + b] += 1;
      counters[2 * B + b] += 1;
    }
  }
  __syncthreads();  // single block: every thread has read *step
  if (threadIdx.x == 0) *step = st + 1;
}

}  // namespace pg

using namespace pg;
#define PG_ST(s) reinterpret_cast<cudaStream_t>(s)
#define PG_RET()     \
  pg_count_launch(1); \
  return cudaGetLastError() == cudaSuccess ? PG_OK : PG_ERR_CUDA

static long long g_launches = 0;
static int g_pdl = 1;
void pg_count_launch(int n) { g_launches += n; }
int pg_pdl_enabled(void) { return g_pdl; }
extern "C" int pg_set_pdl(int on) { g_pdl = on ? 1 : 0; return PG_OK; }
extern "C" long long pg_launch_count(void) { return g_launches; }

extern "C" int pg_abi_version(void) { return 2; }
extern "C" int pg_num_sms(void) { return pg::num_sms(); }

extern "C" int pg_check_device(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return PG_ERR_CUDA;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return PG_ERR_CUDA;
  return major == 10 ? PG_OK : PG_ERR_ARCH;
}

extern "C" int pg_layernorm(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, int rows,
                            int D, float eps, void* stream) {
  if (rows <= 0 || D <= 0 || (D % 4) != 0 || D > 8192) return PG_ERR_ARG;
  const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                    reinterpret_cast<uintptr_t>(y_bf16) | reinterpret_cast<uintptr_t>(y_f32)) & 15) == 0;
  if (rows >= 1024 && (D % 128) == 0 && D <= 2048 && al) {  // prefill: one warp per row, the row stays in registers
    layernorm_warp_kernel<16><<<(rows + 7) / 8, 256, 0, PG_ST(stream)>>>(x, gamma, beta, static_cast<bf16*>(y_bf16), y_f32, rows, D, eps);
    PG_RET();
  }
  layernorm_kernel<<<rows, 256, (D + 33) * sizeof(float), PG_ST(stream)>>>(x, gamma, beta, static_cast<bf16*>(y_bf16), y_f32, D, eps);
  PG_RET();
}

static int g_rmsnorm_early_trigger = 1;
extern "C" int pg_debug_set_rmsnorm_early_trigger(int on) { g_rmsnorm_early_trigger = on; return 0; }

extern "C" int pg_rmsnorm(const float* x, const float* w, void* y_bf16, int rows, int D, float eps, void* stream) {
  if (rows <= 0 || D <= 0 || (D % 4) != 0 || D > 8192) return PG_ERR_ARG;
  if (rows >= 1024 && (D % 128) == 0 && D <= 2048 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y_bf16)) & 15) == 0) {
    rmsnorm_warp_kernel<16><<<(rows + 7) / 8, 256, 0, PG_ST(stream)>>>(x, w, static_cast<bf16*>(y_bf16), rows, D, eps);
    PG_RET();
  }
  return launch_kernel(rmsnorm_kernel, dim3(rows), dim3(256), (D + 33) * sizeof(float), PG_ST(stream), x, w,
                       static_cast<bf16*>(y_bf16), D, eps, g_rmsnorm_early_trigger) == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}

extern "C" int pg_im2col(const float* pixels, void* patches, int B, int C, int H, int W, int P, int Kpad, void* stream) {
  if (B <= 0 || P <= 0 || H % P || W % P || Kpad < C * P * P || (Kpad % 8)) return PG_ERR_ARG;
  const long long total = static_cast<long long>(B) * (H / P) * (W / P) * Kpad;
  im2col_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, PG_ST(stream)>>>(pixels, static_cast<bf16*>(patches), C, H, W, P, Kpad, total);
  PG_RET();
}

extern "C" int pg_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
  if (n <= 0) return PG_ERR_ARG;
  const long long pairs = (n + 1) / 2;
  cast_f32_bf16_kernel<<<static_cast<unsigned>((pairs + 255) / 256), 256, 0, PG_ST(stream)>>>(src, static_cast<bf16*>(dst), n);
  PG_RET();
}

extern "C" int pg_pack_gate_up(const void* gate, const void* up, void* packed, int F, int K, void* stream) {
  if (F <= 0 || K <= 0 || (F % 64) || (K % 8)) return PG_ERR_ARG;
  const long long total = 2ll * F * (K / 8);
  pack_gate_up_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, PG_ST(stream)>>>(
      static_cast<const bf16*>(gate), static_cast<const bf16*>(up), static_cast<bf16*>(packed), F, K / 8);
  PG_RET();
}

extern "C" int pg_merge_embeddings(const long long* input_ids, const long long* attn_mask, const void* embed, const float* img,
                                   float* h, int* pos, int* src_scratch, int* err_flag, int B, int S, int D, int N,
                                   long long image_token, long long pad_token, float text_scale, float img_scale,
                                   void* stream) {
  if (B <= 0 || S <= 0 || D <= 0 || (D % 4)) return PG_ERR_ARG;
  merge_scan_kernel<<<B, 1024, 0, PG_ST(stream)>>>(input_ids, attn_mask, src_scratch, pos, err_flag, nullptr, S, N, image_token, pad_token);
  if (cudaGetLastError() != cudaSuccess) return PG_ERR_CUDA;
  pg_count_launch(1);
  merge_gather_kernel<<<B * S, 256, 0, PG_ST(stream)>>>(input_ids, src_scratch, static_cast<const bf16*>(embed), img, h, S, D, N,
                                                        text_scale, img_scale, 0);
  PG_RET();
}

extern "C" int pg_merge_scan(const long long* input_ids, const long long* attn_mask, int* pos, int* src_scratch, int* dst_row,
                             int* err_flag, int B, int S, int N, long long image_token, long long pad_token, void* stream) {
  if (B <= 0 || S <= 0 || N < 0 || !input_ids || !attn_mask || !pos || !src_scratch || !err_flag) return PG_ERR_ARG;
  merge_scan_kernel<<<B, 1024, 0, PG_ST(stream)>>>(input_ids, attn_mask, src_scratch, pos, err_flag, dst_row, S, N, image_token, pad_token);
  PG_RET();
}

extern "C" int pg_merge_text(const long long* input_ids, const int* src_scratch, const void* embed, float* h, int B, int S, int D,
                             int N, float text_scale, void* stream) {
  if (B <= 0 || S <= 0 || D <= 0 || (D % 4) || !input_ids || !src_scratch || !embed || !h) return PG_ERR_ARG;
  merge_gather_kernel<<<B * S, 256, 0, PG_ST(stream)>>>(input_ids, src_scratch, static_cast<const bf16*>(embed), nullptr, h, S, D, N,
                                                        text_scale, 0.f, 1);
  PG_RET();
}

extern "C" int pg_embed_tokens(const int* tokens, const void* embed, const float* img, float* h, int B, int D, int N,
                               float text_scale, float img_scale, long long pad_token, long long image_token, void* stream) {
  if (B <= 0 || D <= 0 || (D % 4)) return PG_ERR_ARG;
  return launch_kernel(embed_tokens_kernel, dim3(B), dim3(256), 0, PG_ST(stream), tokens, static_cast<const bf16*>(embed), img, h,
                       D, N, text_scale, img_scale, pad_token, image_token) == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}

extern "C" int pg_rope_kv_append(const void* qkv, int qkv_is_f32, const int* pos, void* q_out, void* k_out, void* v_out,
                                 void* k_pages, void* v_pages, const int* page_table, const int* slot_base, int B,
                                 int tokens_per_seq, int Hq, int Hkv, int dh, int page_size, int max_pages,
                                 const float* inv_freq, void* stream) {
  if (B <= 0 || tokens_per_seq <= 0 || (dh % 2)) return PG_ERR_ARG;
  const int T = B * tokens_per_seq;
  if (qkv_is_f32)
    rope_kv_append_kernel<true><<<T, 128, 0, PG_ST(stream)>>>(qkv, pos, static_cast<bf16*>(q_out), static_cast<bf16*>(k_out),
                                                             static_cast<bf16*>(v_out), static_cast<bf16*>(k_pages),
                                                             static_cast<bf16*>(v_pages), page_table, slot_base, tokens_per_seq,
                                                             Hq, Hkv, dh, page_size, max_pages, inv_freq);
  else
    rope_kv_append_kernel<false><<<T, 128, 0, PG_ST(stream)>>>(qkv, pos, static_cast<bf16*>(q_out), static_cast<bf16*>(k_out),
                                                              static_cast<bf16*>(v_out), static_cast<bf16*>(k_pages),
                                                              static_cast<bf16*>(v_pages), page_table, slot_base, tokens_per_seq,
                                                              Hq, Hkv, dh, page_size, max_pages, inv_freq);
  PG_RET();
}

// ---------------------------------------------------------------------------------------------------------
// L2 weight prefetch for the decode chain.  A decode layer streams 220 MB of weights, but for the ~24 us of its attention
// block (norm, q/k/v, attention, o_proj, norm: seven kernel boundaries' worth of latency) the HBM pins idle.  This kernel
// -- launched on a FORKED branch of the step's CUDA graph, so its launch is off the critical path -- asks the bulk-copy
// engine to pull a contiguous weight range into L2 and exits; the GEMM that consumes the range later finds it there.
// One 32-lane CTA per SM, every lane issues cp.async.bulk.prefetch.L2 for `chunk`-byte pieces (no registers, no shared
// memory: co-resides with anything).  Never changes results.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) prefetch_l2_kernel(const char* __restrict__ p, long long bytes, int chunk, int evict_last) {
  const long long n = (bytes + chunk - 1) / chunk;
  for (long long c = static_cast<long long>(blockIdx.x) + static_cast<long long>(threadIdx.x) * gridDim.x; c < n;
       c += static_cast<long long>(gridDim.x) * 32) {
    const long long off = c * chunk;
    const long long len = min(static_cast<long long>(chunk), bytes - off);
    if (evict_last) prefetch_l2_bulk_hint(p + off, static_cast<uint32_t>(len & ~15ll), kEvictLast);
    else prefetch_l2_bulk(p + off, static_cast<uint32_t>(len & ~15ll));
  }
}

extern "C" int pg_prefetch_l2(const void* ptr, long long bytes, int ctas, int evict_last, void* stream) {
  if (ptr == nullptr || bytes < 16 || (reinterpret_cast<uintptr_t>(ptr) & 15) || ctas < 0) return PG_ERR_ARG;
  const int chunk = 16384;
  pg_count_launch(1);
  prefetch_l2_kernel<<<ctas == 0 ? num_sms() : ctas, 32, 0, PG_ST(stream)>>>(static_cast<const char*>(ptr), bytes, chunk, evict_last);
  PG_RET();
}

extern "C" int pg_kv_gather(const void* pages, const int* page_table, void* dense, int B, int len, int Hkv, int dh,
                            int page_size, int max_pages, void* stream) {
  if (B <= 0 || len <= 0) return PG_ERR_ARG;
  const long long total = static_cast<long long>(B) * Hkv * len * dh;
  kv_gather_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, PG_ST(stream)>>>(
      static_cast<const bf16*>(pages), page_table, static_cast<bf16*>(dense), len, Hkv, dh, page_size, max_pages, total);
  PG_RET();
}

extern "C" int pg_advance_decode_slots(const int* next, int* tok_ring, int ring, int* cur_tok, int* counters, const int* kv_limit,
                                       int* step, int B, void* stream) {
  if (B <= 0 || B > 1024 || ring <= 0 || !next || !tok_ring || !cur_tok || !counters || !kv_limit || !step) return PG_ERR_ARG;
  return launch_kernel(advance_decode_slots_kernel, dim3(1), dim3(1024), 0, PG_ST(stream), next, tok_ring, ring, cur_tok,
                       counters, kv_limit, step, B) == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}

extern "C" int pg_advance_decode(const int* next, int* tok_hist, int* cur_tok, int* counters, int n_counters, int* step, int B,
                                 void* stream) {
  if (B <= 0 || B > 1024 || n_counters < 0) return PG_ERR_ARG;
  return launch_kernel(advance_decode_kernel, dim3(1), dim3(1024), 0, PG_ST(stream), next, tok_hist, cur_tok, counters,
                       n_counters, step, B) == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}
