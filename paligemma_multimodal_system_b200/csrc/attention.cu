// Attention kernels (bf16 in/out, fp32 online softmax).
//
//  * attn_prefill_kernel : non-causal flash attention for the SigLIP tower (modeling_siglip.py:96-136, dh = 72) and for
//    the Gemma prefill (modeling_gemma.py:307-339 with the all-zero mask of modeling_paligemma.py:154-156, dh = 256,
//    MQA: the Hq query heads of one token are stacked as consecutive rows against the single KV head, so repeat_kv
//    (modeling_gemma.py:185-196) never materialises).
//  * attn_decode_kernel  : q_len = 1 decode over the paged bf16 KV cache, split over the KV length, K/V tiles staged
//    through shared memory with 16-byte cp.async, warp-shuffle softmax reductions, + a combine kernel.
//
// Round-1 implementation: mma.sync m16n8k16 tensor-core tiles (attention is 1.4 % of the 224-px prefill FLOPs); the
// tcgen05/TMEM version for the 448/896-px configs is the next step for this file.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "paligemma_b200.h"
#include "tmap.cuh"

int pg_attention_prefill_tc(const void* q, const void* k, const void* v, void* o, int B, int H, int rows, int keys, int dh,
                            int group, long long q_bs, long long q_ts, long long q_hs, long long q_head_off, long long kv_bs,
                            long long kv_ts, long long kv_head_off, long long o_bs, long long o_ts, long long o_hs,
                            long long o_head_off, float scale, const int* key_lens, void* stream);  // attention_prefill_tc.cu

namespace pg {

typedef __nv_bfloat16 bf16;

struct AttnPrefillParams {
  const bf16* q;
  const bf16* k;
  const bf16* v;
  bf16* o;
  int rows, keys, group;
  long long q_bs, q_ts, q_hs, q_head_off, kv_bs, kv_ts, kv_head_off, o_bs, o_ts, o_hs, o_head_off;
  float sl2;  // softmax scale * log2(e)
};

template <int DH>
struct AttnCfg {
  static constexpr int DHP = (DH + 15) / 16 * 16;  // padded to the mma K granularity (72 -> 80), pad lanes are zero
  static constexpr int LDS = DHP + 8;              // +16 B row padding: conflict-free ldmatrix
  static constexpr int CHUNKS = DHP / 8;           // 16-byte chunks per (padded) row
  static constexpr int VALID_CHUNKS = DH / 8;
};

// cooperative tile load: `nrows` rows of DH bf16 (row i from src_row(i), nullptr => zero row) into smem [nrows][LDS]
template <int DH, int NTHREADS, typename RowPtr>
PG_DEVINL void load_tile(bf16* smem, int nrows, RowPtr src_row) {
  using C = AttnCfg<DH>;
  for (int idx = threadIdx.x; idx < nrows * C::CHUNKS; idx += NTHREADS) {
    const int r = idx / C::CHUNKS, c = idx % C::CHUNKS;
    const bf16* src = src_row(r);
    const bool valid = (src != nullptr) && (c < C::VALID_CHUNKS);
    // keep the (unused) address in bounds when not valid
    cp_async16(smem_u32(smem + r * C::LDS + c * 8), valid ? static_cast<const void*>(src + c * 8) : static_cast<const void*>(smem), valid);
  }
}

template <int DH, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) attn_prefill_kernel(const AttnPrefillParams p) {
  using C = AttnCfg<DH>;
  constexpr int BLOCK_M = NWARPS * 16;
  constexpr int BLOCK_N = 64;
  constexpr int NT = NWARPS * 32;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Ks = Qs + BLOCK_M * C::LDS;
  bf16* Vs = Ks + 2 * BLOCK_N * C::LDS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BLOCK_M;
  const int h = blockIdx.y, b = blockIdx.z;
  const bf16* qb = p.q + b * p.q_bs + h * p.q_head_off;
  const bf16* kb = p.k + b * p.kv_bs + h * p.kv_head_off;
  const bf16* vb = p.v + b * p.kv_bs + h * p.kv_head_off;

  load_tile<DH, NT>(Qs, BLOCK_M, [&](int r) -> const bf16* {
    const int row = m0 + r;
    return row < p.rows ? qb + (row / p.group) * p.q_ts + (row % p.group) * p.q_hs : nullptr;
  });
  auto load_kv = [&](int tile, int buf) {
    const int n0 = tile * BLOCK_N;
    load_tile<DH, NT>(Ks + buf * BLOCK_N * C::LDS, BLOCK_N, [&](int r) -> const bf16* {
      return (n0 + r) < p.keys ? kb + static_cast<long long>(n0 + r) * p.kv_ts : nullptr;
    });
    load_tile<DH, NT>(Vs + buf * BLOCK_N * C::LDS, BLOCK_N, [&](int r) -> const bf16* {
      return (n0 + r) < p.keys ? vb + static_cast<long long>(n0 + r) * p.kv_ts : nullptr;
    });
  };
  const int ntiles = (p.keys + BLOCK_N - 1) / BLOCK_N;
  load_kv(0, 0);
  cp_async_commit();

  float o[C::DHP / 8][4];
#pragma unroll
  for (int i = 0; i < C::DHP / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  const uint32_t q_addr = smem_u32(Qs + (warp * 16 + (lane & 15)) * C::LDS + (lane >> 4) * 8);

  for (int it = 0; it < ntiles; ++it) {
    const int buf = it & 1;
    if (it + 1 < ntiles) {
      load_kv(it + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    const bf16* Kt = Ks + buf * BLOCK_N * C::LDS;
    const bf16* Vt = Vs + buf * BLOCK_N * C::LDS;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    const uint32_t k_addr = smem_u32(Kt + ((lane & 7) + (lane >> 4) * 8) * C::LDS + ((lane >> 3) & 1) * 8);
#pragma unroll
    for (int ks = 0; ks < C::DHP / 16; ++ks) {
      uint32_t a[4];
      ldmatrix_x4(q_addr + ks * 32, a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(k_addr + (np * 16 * C::LDS + ks * 16) * 2, b0, b1, b2, b3);
        mma_bf16_16816(s[2 * np], a, b0, b1);
        mma_bf16_16816(s[2 * np + 1], a, b2, b3);
      }
    }
    // mask the tail keys of the last tile
    const int n0 = it * BLOCK_N;
    if (n0 + BLOCK_N > p.keys) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = n0 + nt * 8 + (lane & 3) * 2;
        if (key >= p.keys) s[nt][0] = s[nt][2] = -INFINITY;
        if (key + 1 >= p.keys) s[nt][1] = s[nt][3] = -INFINITY;
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float alpha[2], msc[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      alpha[r] = exp2f((m_run[r] - m_safe) * p.sl2);
      msc[r] = m_safe * p.sl2;
      m_run[r] = m_new;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] * p.sl2 - msc[0]);
      s[nt][1] = exp2f(s[nt][1] * p.sl2 - msc[0]);
      s[nt][2] = exp2f(s[nt][2] * p.sl2 - msc[1]);
      s[nt][3] = exp2f(s[nt][3] * p.sl2 - msc[1]);
      rs[0] += s[nt][0] + s[nt][1];
      rs[1] += s[nt][2] + s[nt][3];
    }
    l_run[0] = l_run[0] * alpha[0] + rs[0];
    l_run[1] = l_run[1] * alpha[1] + rs[1];
#pragma unroll
    for (int i = 0; i < C::DHP / 8; ++i) {
      o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
      o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
    }
    const uint32_t v_addr = smem_u32(Vt + ((lane & 7) + ((lane >> 3) & 1) * 8) * C::LDS + (lane >> 4) * 8);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < C::DHP / 16; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(v_addr + (kk * 16 * C::LDS + dp * 16) * 2, b0, b1, b2, b3);
        mma_bf16_16816(o[2 * dp], a, b0, b1);
        mma_bf16_16816(o[2 * dp + 1], a, b2, b3);
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int row0 = m0 + warp * 16 + (lane >> 2);
  bf16* ob = p.o + b * p.o_bs + h * p.o_head_off;
  const long long off0 = (row0 / p.group) * p.o_ts + (row0 % p.group) * p.o_hs;
  const long long off1 = ((row0 + 8) / p.group) * p.o_ts + ((row0 + 8) % p.group) * p.o_hs;
#pragma unroll
  for (int nt = 0; nt < C::DHP / 8; ++nt) {
    const int col = nt * 8 + (lane & 3) * 2;
    if (col < DH) {
      if (row0 < p.rows)
        *reinterpret_cast<uint32_t*>(ob + off0 + col) = pack_bf16(o[nt][0] * inv0, o[nt][1] * inv0);
      if (row0 + 8 < p.rows)
        *reinterpret_cast<uint32_t*>(ob + off1 + col) = pack_bf16(o[nt][2] * inv1, o[nt][3] * inv1);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------------------------
struct AttnDecodeParams {
  const bf16* q;        // [B, Hq*dh]
  const bf16* k_pages;  // [pages, 64, Hkv*dh]
  const bf16* v_pages;
  const int* page_table;  // [B, max_pages]
  const int* kv_len;      // [B]
  float* ws;              // partial O / (m, l)
  int B, Hq, Hkv, max_pages, num_splits;
  float sl2;
};

// workspace layout: o_part [B][Hq][splits][dh], ml_part [B][Hq][splits][2]
template <int DH>
__global__ void __launch_bounds__(128) attn_decode_kernel(const AttnDecodeParams p) {
  using C = AttnCfg<DH>;
  constexpr int BLOCK_N = 64;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);  // [16][LDS]
  bf16* Ks = Qs + 16 * C::LDS;                   // [2][64][LDS]
  bf16* Vs = Ks + 2 * BLOCK_N * C::LDS;
  float* red = reinterpret_cast<float*>(Ks);     // reused after the main loop: [4 warps][16][DH+2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int b = blockIdx.y / p.Hkv, hk = blockIdx.y % p.Hkv;
  const int group = p.Hq / p.Hkv;
  const int len = p.kv_len[b];
  const int n_tiles = (len + BLOCK_N - 1) / BLOCK_N;
  const int tps = (n_tiles + p.num_splits - 1) / p.num_splits;
  const int t_begin = split * tps, t_end = min(n_tiles, t_begin + tps);
  const long long kv_ts = static_cast<long long>(p.Hkv) * DH;

  const bf16* qb = p.q + (static_cast<long long>(b) * p.Hq + hk * group) * DH;
  load_tile<DH, 128>(Qs, 16, [&](int r) -> const bf16* { return r < group ? qb + r * DH : nullptr; });
  auto load_kv = [&](int tile, int buf) {
    const int page = p.page_table[b * p.max_pages + tile];
    const bf16* kb = p.k_pages + static_cast<long long>(page) * BLOCK_N * kv_ts + hk * DH;
    const bf16* vb = p.v_pages + static_cast<long long>(page) * BLOCK_N * kv_ts + hk * DH;
    const int n0 = tile * BLOCK_N;
    load_tile<DH, 128>(Ks + buf * BLOCK_N * C::LDS, BLOCK_N, [&](int r) -> const bf16* { return (n0 + r) < len ? kb + r * kv_ts : nullptr; });
    load_tile<DH, 128>(Vs + buf * BLOCK_N * C::LDS, BLOCK_N, [&](int r) -> const bf16* { return (n0 + r) < len ? vb + r * kv_ts : nullptr; });
  };
  if (t_begin < t_end) load_kv(t_begin, 0);
  cp_async_commit();

  float o[C::DHP / 8][4];
#pragma unroll
  for (int i = 0; i < C::DHP / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const uint32_t q_addr = smem_u32(Qs + (lane & 15) * C::LDS + (lane >> 4) * 8);

  for (int t = t_begin; t < t_end; ++t) {
    const int buf = (t - t_begin) & 1;
    if (t + 1 < t_end) {
      load_kv(t + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    // this warp owns keys [16*warp, 16*warp+16) of the tile
    const bf16* Kt = Ks + (buf * BLOCK_N + warp * 16) * C::LDS;
    const bf16* Vt = Vs + (buf * BLOCK_N + warp * 16) * C::LDS;
    float s[2][4];
    s[0][0] = s[0][1] = s[0][2] = s[0][3] = s[1][0] = s[1][1] = s[1][2] = s[1][3] = 0.f;
    const uint32_t k_addr = smem_u32(Kt + ((lane & 7) + (lane >> 4) * 8) * C::LDS + ((lane >> 3) & 1) * 8);
#pragma unroll
    for (int ks = 0; ks < C::DHP / 16; ++ks) {
      uint32_t a[4], b0, b1, b2, b3;
      ldmatrix_x4(q_addr + ks * 32, a[0], a[1], a[2], a[3]);
      ldmatrix_x4(k_addr + ks * 32, b0, b1, b2, b3);
      mma_bf16_16816(s[0], a, b0, b1);
      mma_bf16_16816(s[1], a, b2, b3);
    }
    const int kbase = t * BLOCK_N + warp * 16;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int key = kbase + nt * 8 + (lane & 3) * 2;
      if (key >= len) s[nt][0] = s[nt][2] = -INFINITY;
      if (key + 1 >= len) s[nt][1] = s[nt][3] = -INFINITY;
    }
    float alpha[2], msc[2], rs[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mx = fmaxf(fmaxf(s[0][2 * r], s[0][2 * r + 1]), fmaxf(s[1][2 * r], s[1][2 * r + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run[r], mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      alpha[r] = exp2f((m_run[r] - m_safe) * p.sl2);
      msc[r] = m_safe * p.sl2;
      m_run[r] = m_new;
      s[0][2 * r] = exp2f(s[0][2 * r] * p.sl2 - msc[r]);
      s[0][2 * r + 1] = exp2f(s[0][2 * r + 1] * p.sl2 - msc[r]);
      s[1][2 * r] = exp2f(s[1][2 * r] * p.sl2 - msc[r]);
      s[1][2 * r + 1] = exp2f(s[1][2 * r + 1] * p.sl2 - msc[r]);
      rs[r] = s[0][2 * r] + s[0][2 * r + 1] + s[1][2 * r] + s[1][2 * r + 1];
      l_run[r] = l_run[r] * alpha[r] + rs[r];
    }
#pragma unroll
    for (int i = 0; i < C::DHP / 8; ++i) {
      o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
      o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
    }
    uint32_t a[4];
    a[0] = pack_bf16(s[0][0], s[0][1]);
    a[1] = pack_bf16(s[0][2], s[0][3]);
    a[2] = pack_bf16(s[1][0], s[1][1]);
    a[3] = pack_bf16(s[1][2], s[1][3]);
    const uint32_t v_addr = smem_u32(Vt + ((lane & 7) + ((lane >> 3) & 1) * 8) * C::LDS + (lane >> 4) * 8);
#pragma unroll
    for (int dp = 0; dp < C::DHP / 16; ++dp) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4_trans(v_addr + dp * 32, b0, b1, b2, b3);
      mma_bf16_16816(o[2 * dp], a, b0, b1);
      mma_bf16_16816(o[2 * dp + 1], a, b2, b3);
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- combine the 4 warps (each saw a disjoint key subset) through shared memory ----
  constexpr int RLD = DH + 2;  // [.., DH] = m (scaled), [.., DH+1] = l
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  {
    const int row = lane >> 2;  // only rows < group (<= 8 of 16) matter; row+8 is padding when group <= 8
    float* dst0 = red + (warp * 16 + row) * RLD;
    float* dst1 = red + (warp * 16 + row + 8) * RLD;
#pragma unroll
    for (int nt = 0; nt < C::DHP / 8; ++nt) {
      const int col = nt * 8 + (lane & 3) * 2;
      if (col < DH) {
        dst0[col] = o[nt][0]; dst0[col + 1] = o[nt][1];
        dst1[col] = o[nt][2]; dst1[col + 1] = o[nt][3];
      }
    }
    if ((lane & 3) == 0) {
      dst0[DH] = m_run[0] * p.sl2; dst0[DH + 1] = l_run[0];
      dst1[DH] = m_run[1] * p.sl2; dst1[DH + 1] = l_run[1];
    }
  }
  __syncthreads();
  float* ws_o = p.ws;
  float* ws_ml = p.ws + static_cast<long long>(p.B) * p.Hq * p.num_splits * DH;
  for (int idx = threadIdx.x; idx < group * DH; idx += 128) {
    const int row = idx / DH, col = idx % DH;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) M = fmaxf(M, red[(w * 16 + row) * RLD + DH]);
    const float Ms = (M == -INFINITY) ? 0.f : M;
    float acc = 0.f, L = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float wgt = exp2f(red[(w * 16 + row) * RLD + DH] - Ms);
      acc += red[(w * 16 + row) * RLD + col] * wgt;
      L += red[(w * 16 + row) * RLD + DH + 1] * wgt;
    }
    const long long hrow = (static_cast<long long>(b) * p.Hq + hk * group + row) * p.num_splits + split;
    ws_o[hrow * DH + col] = acc;
    if (col == 0) {
      ws_ml[hrow * 2] = M;
      ws_ml[hrow * 2 + 1] = L;
    }
  }
}

// out[b, h*dh + c] = sum_s O_s w_s / sum_s L_s w_s,  w_s = 2^(m_s - max m)
__global__ void attn_decode_combine_kernel(const float* ws, bf16* out, int B, int Hq, int dh, int num_splits) {
  const int bh = blockIdx.x;
  const float* ws_o = ws + static_cast<long long>(bh) * num_splits * dh;
  const float* ws_ml = ws + static_cast<long long>(B) * Hq * num_splits * dh + static_cast<long long>(bh) * num_splits * 2;
  float M = -INFINITY;
  for (int s = 0; s < num_splits; ++s) M = fmaxf(M, ws_ml[2 * s]);
  float L = 0.f;
  for (int s = 0; s < num_splits; ++s) L += ws_ml[2 * s + 1] * exp2f(ws_ml[2 * s] - M);
  const float inv = 1.f / L;
  for (int c = threadIdx.x; c < dh; c += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < num_splits; ++s) acc += ws_o[s * dh + c] * exp2f(ws_ml[2 * s] - M);
    out[static_cast<long long>(bh) * dh + c] = __float2bfloat16(acc * inv);
  }
}


// ------------------------------------------------------------------------------------------------------------
// fused decode-step attention: RoPE(q, k_new) + KV append + attention over the paged cache in ONE launch.
// One thread-block CLUSTER per (sequence, kv head): rank r streams a contiguous range of 64-key pages through a
// 3-deep cp.async ring in shared memory (4 warps x 16 keys per page, mma.sync tiles, online softmax), then the ranks
// merge their partial (max, sum, O) through distributed shared memory -- no global workspace, atomics or fences.
// ------------------------------------------------------------------------------------------------------------
struct AttnDecodeFusedParams {
  const float* qkv;       // [B, (Hq+2Hkv)*dh] fp32 raw projections of the new token (pre-RoPE)
  const int* pos;         // [B] position id of the new token
  const int* kv_len;      // [B] cache length INCLUDING the new token (its slot is kv_len-1)
  const float* inv_freq;  // [dh/2]
  bf16* k_pages;          // [pages, 64, Hkv*dh]
  bf16* v_pages;
  const int* page_table;  // [B, max_pages]
  bf16* out;              // [B, Hq*dh]
  int B, Hq, Hkv, max_pages;
  float sl2;
  long long* trace;  // optional profiling stamps (clock64) of CTA 0
};

template <int DH>
__global__ void __launch_bounds__(256) attn_decode_fused_kernel(const __grid_constant__ CUtensorMap tmK,
                                                                const __grid_constant__ CUtensorMap tmV,
                                                                const AttnDecodeFusedParams p) {
  using C = AttnCfg<DH>;
  namespace cg = cooperative_groups;
  constexpr int BLOCK_N = 64;
  constexpr int HALF = DH / 2;
  constexpr int NBUF = 3;
  constexpr int NT = 256;                          // 8 warps: two groups of 4, each group works on its own page
  constexpr int NBOX = DH / 64;                    // 128-byte-wide TMA boxes per K / V page
  constexpr int BOX_BYTES = BLOCK_N * 128;         // 64 rows x 128 B, 128B-swizzled (conflict-free ldmatrix)
  constexpr int KV_BYTES = NBOX * BOX_BYTES;       // one K (or V) page in shared memory
  constexpr int SLOT_BYTES = 2 * KV_BYTES;         // ring slot: K page then V page
  constexpr int RLD = DH + 2;                      // partial row: DH accumulators, m (log2 domain), l
  extern __shared__ __align__(1024) uint8_t smem_dec[];
  uint8_t* ring = smem_dec;                                                // [NBUF][K | V]
  bf16* Qs = reinterpret_cast<bf16*>(smem_dec + NBUF * SLOT_BYTES);        // [16][LDS]
  float* red = reinterpret_cast<float*>(ring);                             // after the loop: [8 warps][16][RLD]
  float* part = red + 8 * 16 * RLD;                                        // this CTA's merged partial [16][RLD]
  static_assert(8 * 16 * RLD * 4 + 16 * RLD * 4 <= NBUF * SLOT_BYTES, "reduction staging must fit in the page ring");
  __shared__ bf16 new_k[DH], new_v[DH];
  __shared__ __align__(8) uint64_t bars[NBUF];
  const uint32_t ring_u32 = smem_u32(ring);
  if ((ring_u32 & 1023u) != 0) __trap();

  cg::cluster_group cluster = cg::this_cluster();
  const int CS = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = warp >> 2, wq = warp & 3;  // warp group (which page of a pair), warp within the group (which 16 keys)
  const int seq = blockIdx.x / CS;
  const int b = seq / p.Hkv, hk = seq % p.Hkv;
  const int group = p.Hq / p.Hkv;
  const bool tr = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (tr) p.trace[0] = clock64();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < NBUF; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_fence_init();
  }
  griddep_wait();
  if (tr) p.trace[1] = clock64();
  if (threadIdx.x == 0) griddep_launch_dependents();
  __syncthreads();
  const int len = p.kv_len[b];
  const int n_tiles = (len + BLOCK_N - 1) / BLOCK_N;
  const int tps = (n_tiles + CS - 1) / CS;
  const int t_begin = min(n_tiles, rank * tps), t_end = min(n_tiles, t_begin + tps);
  const int n_my = t_end - t_begin;
  const int new_slot = len - 1;
  const int new_tile = new_slot / BLOCK_N;
  const bool owns_new = (new_tile >= t_begin && new_tile < t_end);
  const long long kv_ts = static_cast<long long>(p.Hkv) * DH;
  const int* ptab = p.page_table + b * p.max_pages;

  // One page = NBOX TMA boxes for K and NBOX for V (64 rows x 128 B each, hardware 128B swizzle), all on the slot's
  // mbarrier.  Rows beyond kv_len hold zeros / stale finite values from the page pool: their scores are masked and
  // their probabilities are exactly 0.  The new token's row is patched in from registers below.
  // (issued by thread 128, which has no RoPE work, so that the page loads overlap the query staging)
  auto load_kv = [&](int tile, int slot) {
    if (threadIdx.x == 128) {
      const int page = __ldg(ptab + tile);
      const uint32_t bar = smem_u32(&bars[slot]);
      const uint32_t dst = ring_u32 + slot * SLOT_BYTES;
      mbar_expect_tx(bar, SLOT_BYTES);
#pragma unroll
      for (int bx = 0; bx < NBOX; ++bx) {
        tma_load_2d(dst + bx * BOX_BYTES, &tmK, bar, hk * DH + bx * 64, page * BLOCK_N, kEvictFirst);
        tma_load_2d(dst + KV_BYTES + bx * BOX_BYTES, &tmV, bar, hk * DH + bx * 64, page * BLOCK_N, kEvictFirst);
      }
    }
  };
  // byte offset of element (row r, column c) inside a swizzled K (or V) page
  auto swz = [&](int r, int c) -> int { return (c >> 6) * BOX_BYTES + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2; };
  auto patch_new_row = [&](int slot) {
    uint8_t* Kb = ring + slot * SLOT_BYTES;
    const int r = new_slot - new_tile * BLOCK_N;
    for (int k = threadIdx.x; k < DH; k += NT) {
      *reinterpret_cast<bf16*>(Kb + swz(r, k)) = new_k[k];
      *reinterpret_cast<bf16*>(Kb + KV_BYTES + swz(r, k)) = new_v[k];
    }
  };
  if (threadIdx.x == 128) {  // everything this rank can hold goes in flight now; page ids are fetched together first
    int pages[NBUF];
#pragma unroll
    for (int i = 0; i < NBUF; ++i) pages[i] = (i < n_my) ? __ldg(ptab + t_begin + i) : 0;
#pragma unroll
    for (int i = 0; i < NBUF; ++i) {
      if (i < n_my) {
        const uint32_t bar = smem_u32(&bars[i]);
        const uint32_t dst = ring_u32 + i * SLOT_BYTES;
        mbar_expect_tx(bar, SLOT_BYTES);
#pragma unroll
        for (int bx = 0; bx < NBOX; ++bx) {
          tma_load_2d(dst + bx * BOX_BYTES, &tmK, bar, hk * DH + bx * 64, pages[i] * BLOCK_N, kEvictFirst);
          tma_load_2d(dst + KV_BYTES + bx * BOX_BYTES, &tmV, bar, hk * DH + bx * 64, pages[i] * BLOCK_N, kEvictFirst);
        }
      }
    }
  }

  // RoPE (rotate-half, modeling_gemma.py:138-151) on the query heads of this group (+ the new key when owned).
  // All global loads of a thread are issued back to back (independent), then rotated: one L2 round trip, not 20.
  const int W = (p.Hq + 2 * p.Hkv) * DH;
  const float* __restrict__ row = p.qkv + static_cast<long long>(b) * W;
  const float posf = static_cast<float>(__ldg(p.pos + b));
  const int new_page = owns_new ? __ldg(ptab + new_tile) : 0;
  for (int i = threadIdx.x; i < HALF; i += NT) {
    const float freq = __ldg(p.inv_freq + i);
    float x1[16], x2[16], kx1 = 0.f, kx2 = 0.f, vx1 = 0.f, vx2 = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      if (g < group) {
        const float* qh = row + (hk * group + g) * DH;
        x1[g] = __ldcg(qh + i);
        x2[g] = __ldcg(qh + i + HALF);
      }
    }
    if (owns_new) {
      const float* kh = row + (p.Hq + hk) * DH;
      const float* vh = row + (p.Hq + p.Hkv + hk) * DH;
      kx1 = __ldcg(kh + i); kx2 = __ldcg(kh + i + HALF);
      vx1 = __ldcg(vh + i); vx2 = __ldcg(vh + i + HALF);
    }
    float sn, cs;
    sincosf(posf * freq, &sn, &cs);
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      if (g < group) {
        Qs[g * C::LDS + i] = __float2bfloat16(x1[g] * cs - x2[g] * sn);
        Qs[g * C::LDS + i + HALF] = __float2bfloat16(x2[g] * cs + x1[g] * sn);
      }
    }
    if (owns_new) {
      const bf16 k1 = __float2bfloat16(kx1 * cs - kx2 * sn), k2 = __float2bfloat16(kx2 * cs + kx1 * sn);
      const bf16 v1 = __float2bfloat16(vx1), v2 = __float2bfloat16(vx2);
      const int page = new_page;
      bf16* kb = p.k_pages + (static_cast<long long>(page) * BLOCK_N + (new_slot - new_tile * BLOCK_N)) * kv_ts + hk * DH;
      bf16* vb = p.v_pages + (static_cast<long long>(page) * BLOCK_N + (new_slot - new_tile * BLOCK_N)) * kv_ts + hk * DH;
      kb[i] = k1; kb[i + HALF] = k2;   // KVCache.update (modeling_gemma.py:18-57)
      vb[i] = v1; vb[i + HALF] = v2;
      new_k[i] = k1; new_k[i + HALF] = k2;  // staged copy, patched into the shared-memory page below
      new_v[i] = v1; new_v[i + HALF] = v2;
    }
  }
  // zero the padding rows / pad columns of Q
  for (int idx = threadIdx.x; idx < 16 * C::DHP; idx += NT) {
    const int r = idx / C::DHP, cc = idx % C::DHP;
    if (r >= group || cc >= DH) Qs[r * C::LDS + cc] = __float2bfloat16(0.f);
  }
  __syncthreads();  // Q, new_k / new_v staged
  if (tr) p.trace[2] = clock64();

  float o[C::DHP / 8][4];
#pragma unroll
  for (int i = 0; i < C::DHP / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const uint32_t q_addr = smem_u32(Qs + (lane & 15) * C::LDS + (lane >> 4) * 8);

  // pages are consumed in pairs: warp group 0 takes page 2j, group 1 takes page 2j+1 (both groups stay in lock step
  // through the CTA barriers so that ring slots can be recycled)
  for (int i0 = 0; i0 < n_my; i0 += 2) {
    const int i = i0 + wg;
    const bool active = i < n_my;
    if (active) {
      const int t = t_begin + i;
      const int slot = i % NBUF;
      mbar_wait(smem_u32(&bars[slot]), (i / NBUF) & 1);
      if (tr && i0 == 0) p.trace[3] = clock64();
      if (t == new_tile) {  // the TMA has landed: overwrite the new token's row (only this warp group reads this page)
        const int r = new_slot - new_tile * BLOCK_N;
        uint8_t* Kb = ring + slot * SLOT_BYTES;
        for (int k = (warp & 3) * 32 + lane; k < DH; k += 128) {
          *reinterpret_cast<bf16*>(Kb + swz(r, k)) = new_k[k];
          *reinterpret_cast<bf16*>(Kb + KV_BYTES + swz(r, k)) = new_v[k];
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");  // the 4 warps of this group
      }
      const uint32_t k_base = ring_u32 + slot * SLOT_BYTES;
      const uint32_t v_base = k_base + KV_BYTES;
      float s[2][4], s2[2][4];  // two independent accumulator sets (even / odd k-steps): 4 MMA chains in flight
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[a][e] = s2[a][e] = 0.f;
      // ldmatrix row addresses in the swizzled page: row kr, 16-byte chunk index XOR (kr & 7)
      const int kr = wq * 16 + (lane & 7) + (lane >> 4) * 8;
      const uint32_t k_row = k_base + kr * 128;
      const int k_sub = (lane >> 3) & 1;  // which 8-column half of the 16-wide k-step
#pragma unroll
      for (int ks = 0; ks < C::DHP / 16; ks += 2) {
        uint32_t a[4], b0, b1, b2, b3;
        ldmatrix_x4(q_addr + ks * 32, a[0], a[1], a[2], a[3]);
        ldmatrix_x4(k_row + (ks >> 2) * BOX_BYTES + (((((ks & 3) << 1) + k_sub) ^ (kr & 7)) << 4), b0, b1, b2, b3);
        mma_bf16_16816(s[0], a, b0, b1);
        mma_bf16_16816(s[1], a, b2, b3);
        if (ks + 1 < C::DHP / 16) {
          uint32_t c4[4], d0, d1, d2, d3;
          ldmatrix_x4(q_addr + (ks + 1) * 32, c4[0], c4[1], c4[2], c4[3]);
          ldmatrix_x4(k_row + ((ks + 1) >> 2) * BOX_BYTES + ((((((ks + 1) & 3) << 1) + k_sub) ^ (kr & 7)) << 4), d0, d1, d2, d3);
          mma_bf16_16816(s2[0], c4, d0, d1);
          mma_bf16_16816(s2[1], c4, d2, d3);
        }
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[a][e] += s2[a][e];
      const int kbase = t * BLOCK_N + wq * 16;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int key = kbase + nt * 8 + (lane & 3) * 2;
        if (key >= len) s[nt][0] = s[nt][2] = -INFINITY;
        if (key + 1 >= len) s[nt][1] = s[nt][3] = -INFINITY;
      }
      float alpha[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float mx = fmaxf(fmaxf(s[0][2 * r], s[0][2 * r + 1]), fmaxf(s[1][2 * r], s[1][2 * r + 1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float m_new = fmaxf(m_run[r], mx);
        const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
        alpha[r] = exp2f((m_run[r] - m_safe) * p.sl2);
        const float msc = m_safe * p.sl2;
        m_run[r] = m_new;
        s[0][2 * r] = exp2f(s[0][2 * r] * p.sl2 - msc);
        s[0][2 * r + 1] = exp2f(s[0][2 * r + 1] * p.sl2 - msc);
        s[1][2 * r] = exp2f(s[1][2 * r] * p.sl2 - msc);
        s[1][2 * r + 1] = exp2f(s[1][2 * r + 1] * p.sl2 - msc);
        l_run[r] = l_run[r] * alpha[r] + s[0][2 * r] + s[0][2 * r + 1] + s[1][2 * r] + s[1][2 * r + 1];
      }
      if (i0 > 0) {  // (first pair: the accumulators are still zero)
#pragma unroll
        for (int k = 0; k < C::DHP / 8; ++k) {
          o[k][0] *= alpha[0]; o[k][1] *= alpha[0];
          o[k][2] *= alpha[1]; o[k][3] *= alpha[1];
        }
      }
      uint32_t a[4];
      a[0] = pack_bf16(s[0][0], s[0][1]);
      a[1] = pack_bf16(s[0][2], s[0][3]);
      a[2] = pack_bf16(s[1][0], s[1][1]);
      a[3] = pack_bf16(s[1][2], s[1][3]);
      const int vr = wq * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const uint32_t v_row = v_base + vr * 128;
      const int v_sub = lane >> 4;
#pragma unroll
      for (int dp = 0; dp < C::DHP / 16; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(v_row + (dp >> 2) * BOX_BYTES + (((((dp & 3) << 1) + v_sub) ^ (vr & 7)) << 4), b0, b1, b2, b3);
        mma_bf16_16816(o[2 * dp], a, b0, b1);
        mma_bf16_16816(o[2 * dp + 1], a, b2, b3);
      }
    }
    __syncthreads();  // both pages of the pair are consumed
    // refill the two slots just freed (long contexts): pages i0 + NBUF, i0 + 1 + NBUF
    if (i0 + NBUF < n_my) {
      fence_proxy_async_smem();
      for (int k = 0; k < 2; ++k) {
        const int inext = i0 + k + NBUF;
        if (i0 + k < n_my && inext < n_my) load_kv(t_begin + inext, inext % NBUF);
      }
    }
  }
  __syncthreads();

  if (tr) p.trace[4] = clock64();
  // ---- merge the 8 warps (disjoint key subsets) through shared memory ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  {
    const int r0 = lane >> 2;
    float* dst0 = red + (warp * 16 + r0) * RLD;
    float* dst1 = red + (warp * 16 + r0 + 8) * RLD;
    const bool hi_rows = group > 8;  // rows 8..15 are padding unless the GQA group is larger than 8
#pragma unroll
    for (int nt = 0; nt < C::DHP / 8; ++nt) {
      const int col = nt * 8 + (lane & 3) * 2;
      if (col < DH) {
        dst0[col] = o[nt][0]; dst0[col + 1] = o[nt][1];
        if (hi_rows) { dst1[col] = o[nt][2]; dst1[col + 1] = o[nt][3]; }
      }
    }
    if ((lane & 3) == 0) {
      dst0[DH] = m_run[0] * p.sl2; dst0[DH + 1] = l_run[0];
      if (hi_rows) { dst1[DH] = m_run[1] * p.sl2; dst1[DH + 1] = l_run[1]; }
    }
  }
  __syncthreads();
  const long long hq0 = static_cast<long long>(b) * p.Hq + hk * group;  // first query head of this group
  __shared__ float s_w[8][16];   // weight of warp w's partial for row r (already divided by the row sum when CS == 1)
  __shared__ float s_ML[16][2];  // this CTA's merged (max, sum) per row
  if (threadIdx.x < group) {
    const int r = threadIdx.x;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) M = fmaxf(M, red[(w * 16 + r) * RLD + DH]);
    const float Ms = (M == -INFINITY) ? 0.f : M;
    float Lsum = 0.f, wv[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      wv[w] = exp2f(red[(w * 16 + r) * RLD + DH] - Ms);
      Lsum += red[(w * 16 + r) * RLD + DH + 1] * wv[w];
    }
    const float norm = (CS == 1) ? 1.f / Lsum : 1.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s_w[w][r] = wv[w] * norm;
    s_ML[r][0] = M;
    s_ML[r][1] = Lsum;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < group * DH; idx += NT) {
    const int r = idx / DH, col = idx % DH;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) acc += red[(w * 16 + r) * RLD + col] * s_w[w][r];
    if (CS == 1) p.out[(hq0 + r) * DH + col] = __float2bfloat16(acc);
    else part[r * RLD + col] = acc;
  }
  if (CS > 1 && threadIdx.x < group) {
    part[threadIdx.x * RLD + DH] = s_ML[threadIdx.x][0];
    part[threadIdx.x * RLD + DH + 1] = s_ML[threadIdx.x][1];
  }
  if (CS == 1) return;

  if (tr) p.trace[5] = clock64();
  // ---- merge the ranks through distributed shared memory; rank q finalises columns [q*DH/CS, (q+1)*DH/CS) ----
  cluster.sync();
  if (tr) p.trace[6] = clock64();
  {
    __shared__ float s_rw[16][8];  // weight (incl. 1 / row sum) of rank q's partial for row r
    if (threadIdx.x < group) {
      const int r = threadIdx.x;
      float mv[8], lv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {  // independent remote loads, issued back to back
        if (q < CS) {
          const float* rp = cluster.map_shared_rank(part, q);
          mv[q] = rp[r * RLD + DH];
          lv[q] = rp[r * RLD + DH + 1];
        }
      }
      float M = -INFINITY;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < CS) M = fmaxf(M, mv[q]);
      float Lsum = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < CS) Lsum += lv[q] * exp2f(mv[q] - M);  // empty ranks carry m = -inf, l = 0
      const float inv = 1.f / Lsum;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < CS) s_rw[r][q] = exp2f(mv[q] - M) * inv;
    }
    __syncthreads();
    const int cols_per = (DH + CS - 1) / CS;
    const int c_lo = rank * cols_per, c_hi = min(DH, c_lo + cols_per);
    const int ncol = max(0, c_hi - c_lo);
    for (int idx = threadIdx.x; idx < group * ncol; idx += NT) {
      const int r = idx / ncol, col = c_lo + idx % ncol;
      float av[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < CS) av[q] = cluster.map_shared_rank(part, q)[r * RLD + col];
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < CS) acc += av[q] * s_rw[r][q];
      p.out[(hq0 + r) * DH + col] = __float2bfloat16(acc);
    }
  }
  cluster.sync();  // shared memory must stay alive until every rank has read it
  if (tr) p.trace[7] = clock64();
}

template <int DH>
static int launch_decode_fused(const AttnDecodeFusedParams& p, int num_pages, int cluster_size, cudaStream_t st) {
  using C = AttnCfg<DH>;
  constexpr int smem = 3 * 2 * (DH / 64) * 64 * 128 + 16 * C::LDS * 2;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_decode_fused_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();
      return PG_ERR_CUDA;
    }
    configured = true;
  }
  // the page pools viewed as 2D [num_pages * 64 rows, Hkv * dh columns]; one box = 64 rows x 64 columns (128 B)
  CUtensorMap tmK, tmV;
  int rc;
  const long long cols = static_cast<long long>(p.Hkv) * DH;
  if ((rc = make_tmap_2d(&tmK, p.k_pages, static_cast<long long>(num_pages) * 64, cols, cols, 64)) != PG_OK) return rc;
  if ((rc = make_tmap_2d(&tmV, p.v_pages, static_cast<long long>(num_pages) * 64, cols, cols, 64)) != PG_OK) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(p.B * p.Hkv * cluster_size));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pg_pdl_enabled() ? 2 : 1;
  pg_count_launch(1);
  return cudaLaunchKernelEx(&cfg, attn_decode_fused_kernel<DH>, tmK, tmV, p) == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}

template <int DH, int NWARPS>
static int launch_prefill(const AttnPrefillParams& p, int B, int H, cudaStream_t st) {
  using C = AttnCfg<DH>;
  constexpr int BLOCK_M = NWARPS * 16;
  constexpr int smem = (BLOCK_M + 4 * 64) * C::LDS * 2;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_prefill_kernel<DH, NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return PG_ERR_CUDA;
    configured = true;
  }
  dim3 grid((p.rows + BLOCK_M - 1) / BLOCK_M, H, B);
  attn_prefill_kernel<DH, NWARPS><<<grid, NWARPS * 32, smem, st>>>(p);
  pg_count_launch(1);
  return cudaGetLastError() == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}

template <int DH>
static int launch_decode(const AttnDecodeParams& p, bf16* out, cudaStream_t st) {
  using C = AttnCfg<DH>;
  constexpr int smem_main = (16 + 4 * 64) * C::LDS * 2;
  constexpr int smem_red = 16 * C::LDS * 2 + 4 * 16 * (DH + 2) * 4;
  constexpr int smem = smem_main > smem_red ? smem_main : smem_red;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_decode_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return PG_ERR_CUDA;
    configured = true;
  }
  dim3 grid(p.num_splits, p.B * p.Hkv);
  attn_decode_kernel<DH><<<grid, 128, smem, st>>>(p);
  pg_count_launch(2);
  if (cudaGetLastError() != cudaSuccess) return PG_ERR_CUDA;
  attn_decode_combine_kernel<<<p.B * p.Hq, DH >= 128 ? 128 : 64, 0, st>>>(p.ws, out, p.B, p.Hq, DH, p.num_splits);
  return cudaGetLastError() == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}

}  // namespace pg

using namespace pg;

extern "C" int pg_attention_prefill(const void* q, const void* k, const void* v, void* o, int B, int H, int rows, int keys,
                                    int dh, int group, long long q_bs, long long q_ts, long long q_hs, long long q_head_off,
                                    long long kv_bs, long long kv_ts, long long kv_head_off, long long o_bs, long long o_ts,
                                    long long o_hs, long long o_head_off, float scale, void* stream) {
  if (B <= 0 || H <= 0 || rows <= 0 || keys <= 0 || group <= 0) return PG_ERR_ARG;
  if ((q_bs | q_ts | q_hs | q_head_off | kv_bs | kv_ts | kv_head_off) & 7) return PG_ERR_ARG;  // 16 B cp.async granularity
  if ((o_bs | o_ts | o_hs | o_head_off) & 1) return PG_ERR_ARG;
  if (B > 65535 || H > 65535) return PG_ERR_ARG;
  {
    // tcgen05 / TMEM kernel (attention_prefill_tc.cu) for every shape its TMA maps can express; the mma.sync kernel below
    // only serves the rest (PG_ATTN_PREFILL_MMA=1 forces it: A/B profiling)
    static const bool force_mma = getenv("PG_ATTN_PREFILL_MMA") != nullptr;
    if (!force_mma) {
      const int rc = pg_attention_prefill_tc(q, k, v, o, B, H, rows, keys, dh, group, q_bs, q_ts, q_hs, q_head_off, kv_bs, kv_ts,
                                             kv_head_off, o_bs, o_ts, o_hs, o_head_off, scale, nullptr, stream);
      if (rc <= 0) return rc;
    }
  }
  AttnPrefillParams p;
  p.q = static_cast<const bf16*>(q); p.k = static_cast<const bf16*>(k); p.v = static_cast<const bf16*>(v);
  p.o = static_cast<bf16*>(o);
  p.rows = rows; p.keys = keys; p.group = group;
  p.q_bs = q_bs; p.q_ts = q_ts; p.q_hs = q_hs; p.q_head_off = q_head_off;
  p.kv_bs = kv_bs; p.kv_ts = kv_ts; p.kv_head_off = kv_head_off;
  p.o_bs = o_bs; p.o_ts = o_ts; p.o_hs = o_hs; p.o_head_off = o_head_off;
  p.sl2 = scale * 1.4426950408889634f;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dh) {
    case 64: return launch_prefill<64, 8>(p, B, H, st);
    case 72: return launch_prefill<72, 8>(p, B, H, st);
    case 256: return launch_prefill<256, 8>(p, B, H, st);
    default: return PG_ERR_ARG;
  }
}

extern "C" int pg_attention_prefill_varlen(const void* q, const void* k, const void* v, void* o, const int* key_lens, int B, int H,
                                           int rows, int keys, int dh, int group, long long q_bs, long long q_ts, long long q_hs,
                                           long long q_head_off, long long kv_bs, long long kv_ts, long long kv_head_off,
                                           long long o_bs, long long o_ts, long long o_hs, long long o_head_off, float scale,
                                           void* stream) {
  if (B <= 0 || H <= 0 || rows <= 0 || keys <= 0 || group <= 0 || key_lens == nullptr) return PG_ERR_ARG;
  if (B > 65535 || H > 65535) return PG_ERR_ARG;
  // ragged key counts exist on the tcgen05 kernel only: a shape its TMA maps cannot express is an argument error here
  const int rc = pg_attention_prefill_tc(q, k, v, o, B, H, rows, keys, dh, group, q_bs, q_ts, q_hs, q_head_off, kv_bs, kv_ts,
                                         kv_head_off, o_bs, o_ts, o_hs, o_head_off, scale, key_lens, stream);
  return rc > 0 ? PG_ERR_ARG : rc;
}

extern "C" long long pg_attention_decode_workspace_floats(int B, int Hq, int dh, int num_splits) {
  return static_cast<long long>(B) * Hq * num_splits * (dh + 2);
}

extern "C" int pg_attention_decode(const void* q, const void* k_pages, const void* v_pages, const int* page_table,
                                   const int* kv_len, void* out, float* workspace, int B, int Hq, int Hkv, int dh,
                                   int page_size, int max_pages, int num_splits, float scale, void* stream) {
  if (B <= 0 || Hq <= 0 || Hkv <= 0 || Hq % Hkv != 0 || Hq / Hkv > 16 || page_size != 64 || num_splits <= 0) return PG_ERR_ARG;
  AttnDecodeParams p;
  p.q = static_cast<const bf16*>(q);
  p.k_pages = static_cast<const bf16*>(k_pages);
  p.v_pages = static_cast<const bf16*>(v_pages);
  p.page_table = page_table; p.kv_len = kv_len; p.ws = workspace;
  p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.max_pages = max_pages; p.num_splits = num_splits;
  p.sl2 = scale * 1.4426950408889634f;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dh) {
    case 64: return launch_decode<64>(p, static_cast<bf16*>(out), st);
    case 256: return launch_decode<256>(p, static_cast<bf16*>(out), st);
    default: return PG_ERR_ARG;
  }
}

int pg_attention_decode_v3(const float* qkv, const int* pos, const int* kv_len, const float* inv_freq, void* k_pages,
                           void* v_pages, const int* page_table, void* out, int B, int Hq, int Hkv, int dh, int num_pages,
                           int max_pages, float sl2, int cluster_size, long long* trace, void* stream);  // attention_decode.cu

static long long* g_attn_trace = nullptr;
static int g_attn_trace_idx = 0;
extern "C" int pg_debug_set_attn_trace(long long* p) { g_attn_trace = p; g_attn_trace_idx = 0; return 0; }

extern "C" int pg_attention_decode_fused(const float* qkv, const int* pos, const int* kv_len, const float* inv_freq,
                                         void* k_pages, void* v_pages, const int* page_table, void* out, int B, int Hq,
                                         int Hkv, int dh, int page_size, int num_pages, int max_pages, float scale,
                                         void* stream) {
  if (B <= 0 || Hq <= 0 || Hkv <= 0 || Hq % Hkv != 0 || Hq / Hkv > 16 || page_size != 64 || max_pages <= 0 || num_pages <= 0)
    return PG_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(k_pages) & 15) || (reinterpret_cast<uintptr_t>(v_pages) & 15)) return PG_ERR_ARG;
  AttnDecodeFusedParams p;
  p.qkv = qkv; p.pos = pos; p.kv_len = kv_len; p.inv_freq = inv_freq;
  p.k_pages = static_cast<bf16*>(k_pages); p.v_pages = static_cast<bf16*>(v_pages);
  p.page_table = page_table; p.out = static_cast<bf16*>(out);
  p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.max_pages = max_pages;
  p.sl2 = scale * 1.4426950408889634f;
  p.trace = g_attn_trace ? g_attn_trace + 8 * (g_attn_trace_idx++ % 32) : nullptr;
  // cluster size: as many CTAs as fit in ONE wave (the 3-deep page ring allows one CTA per SM), at most one page per
  // rank, at most 8 (portable cluster limit)
  int cs = 148 / (B * Hkv);
  if (cs > max_pages) cs = max_pages;
  cs = cs >= 8 ? 8 : cs >= 4 ? 4 : cs >= 2 ? 2 : 1;
  static const bool use_v2 = getenv("PG_ATTN_V2") != nullptr;  // A/B switch: the second-generation kernel below
  if (Hq / Hkv <= 8 && !use_v2)
    return pg_attention_decode_v3(qkv, pos, kv_len, inv_freq, k_pages, v_pages, page_table, out, B, Hq, Hkv, dh, num_pages,
                                  max_pages, p.sl2, cs, p.trace, stream);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dh) {
    case 64: return launch_decode_fused<64>(p, num_pages, cs, st);
    case 256: return launch_decode_fused<256>(p, num_pages, cs, st);
    default: return PG_ERR_ARG;
  }
}
