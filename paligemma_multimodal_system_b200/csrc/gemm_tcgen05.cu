// tcgen05 / TMEM / TMA GEMM for every Linear on the PaliGemma hot path.
//
//   acc[t, f] = sum_k X[t, k] * W[f, k]          X: activations [tokens, K] bf16, W: nn.Linear weight [features, K] bf16
//
// Replaces the reference's nn.Linear / nn.Conv2d(im2col) call sites:
//   modeling_siglip.py:59-62,71-75,156,177-185,258-263   modeling_paligemma.py:57,64
//   modeling_gemma.py:205-218,255-259,274-278,356,484,523
//
// One persistent, warp-specialised kernel (warp 0 = TMA producer, warp 1 = tcgen05.mma issuer + TMEM owner,
// warps 2..5 = epilogue).  Operands are staged by TMA into 128B-swizzled shared memory; the fp32 accumulator
// lives in TMEM (double buffered, so the epilogue of tile i overlaps the main loop of tile i+1).
//
//   SWAP = false (prefill, many tokens):  UMMA M = 128 tokens,   N = BN features
//   SWAP = true  (decode, tokens <= 128): UMMA M = 128 features, N = BN tokens  (weight streaming, HBM bound),
//                                         optional split-K with fp32 red.global.add into the output
#include <cmath>

#include "common.cuh"
#include "paligemma_b200.h"
#include "tmap.cuh"

namespace pg {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KB
constexpr int NUM_THREADS = 192;
constexpr int ACC_STAGES = 2;

__host__ __device__ constexpr int b_tile_bytes(int BN) { return BN * BK * 2; }
__host__ __device__ constexpr int stage_bytes(int BN) { return A_TILE_BYTES + b_tile_bytes(BN); }
// swap: GeGLU exchange area; token-major: eight warp-private 4 KB transposition tiles (coalescing epilogues)
__host__ __device__ constexpr int xch_bytes(int BN, bool swap) { return swap ? 64 * BN * 4 : 8 * 4096; }
// decode (SWAP, BN <= 64): two CTAs per SM (115712 B each) so that, with programmatic dependent launch, the next
// kernel's CTAs become resident and prefetch their weights while this kernel drains; otherwise one CTA with <= 200 KB
__host__ __device__ constexpr bool two_per_sm(int BN, bool swap) { return swap && BN <= 64; }
__host__ __device__ constexpr int num_stages(int BN, bool swap) {
  int budget = two_per_sm(BN, swap) ? 115712 - 256 : (swap ? 200 * 1024 : 232448 - 256);
  int s = (budget - xch_bytes(BN, swap)) / stage_bytes(BN);
  return s > 8 ? 8 : s;
}
__host__ __device__ constexpr int tmem_cols(int BN) {
  int c = ACC_STAGES * BN;
  return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512;
}
__host__ __device__ constexpr int smem_bytes(int BN, bool swap) {
  return num_stages(BN, swap) * stage_bytes(BN) + xch_bytes(BN, swap) + 256 /*barriers*/;
}

struct GemmArgs {
  int tokens, features, K;
  int split_k;
  int mode, act_gelu;
  float scale;
  void* out;
  long long ldo;
  const float* bias;
  const float* resid;
  long long ldr;
  const int* out_row_map;  // token-major fp32 epilogues: token t is written to output row out_row_map[t] (projector -> merged rows)
  int resid_mod;  // > 0 (token-major kernels): the residual row of token t is t % resid_mod (a [resid_mod, features] table
                  // broadcast over the batch: the position embeddings added to the patch embeddings, modeling_siglip.py:289-298)
  int n_fast;  // tile raster order (decode_tile)
  int f32_coalesced;  // token-major fp32 epilogue through the shared-memory transposition (alignment checked by the host)
  long long* trace;  // optional profiling stamps (clock64) written by CTA 0
  // ---- decode-step chores of the SWAP kernels ----
  float* zero_buf;  // zero-filled after the dependency wait (the split-K accumulator of a LATER kernel of the chain)
  long long zero_count;
  // ---- ROPE kernels (token-major q/k/v projection of the prefill, one head per N tile: BN == head_dim) ----
  const int* rope_pos;         // [tokens] position ids
  const float* rope_inv_freq;  // [head_dim / 2]
  int rope_hq, rope_hkv;       // query / key-value heads: tile n < hq -> q, < hq + hkv -> k, else v
  __nv_bfloat16* k_pages;      // paged KV cache of this layer (may be null: no cache)
  __nv_bfloat16* v_pages;
  const int* page_table;       // [sequences, max_pages]
  const int* slot_base;        // [sequences] first cache slot of the sequence's tokens
  int tokens_per_seq, max_pages;
  // lm_head (PG_EPI_F32, swap): softmax statistics of every 32-row vocabulary segment, for the sampler that follows
  float2* stats;       // [segments][stats_ld]: (max logit of the segment, sum exp2((x - max) * stat_c)) per token
  long long stats_ld;  // tokens per segment row (>= tokens); 4 * ceil(features / 128) segment rows
  float stat_c;        // inv_temperature * log2(e)
};

struct TileInfo {
  int m_blk, n_blk, kb0, kb1;
};

// Tile order of the persistent grid.  m-fast: concurrently running CTAs share one weight (B) tile and walk the token
// (A) tiles, right when A fits in L2 or the weight matrix is huge (gate||up).  n-fast: concurrently running CTAs cover
// every feature tile of a few token tiles, so each A tile is fetched from DRAM once and the (small) weight matrix stays
// L2 resident: down_proj / fc2 re-read their 0.1-0.5 GB activation matrix once per feature tile otherwise (ncu: 4.1 GB
// of DRAM reads for a 0.6 GB problem).
// banded (n_fast = 2 + band): both operands exceed what L2 keeps -- gate||up at prefill sizes: 134 MB of activations AND 134 MB
// of weights; m-fast re-read the activation matrix once per feature tile (ncu: 10.9 GB of DRAM reads for a 0.27 GB problem).
// The token tiles are walked in bands of `band` tiles (~32 MB of activations, L2 resident while every feature tile passes),
// so the weights are streamed once per band and the activations once.
PG_DEVINL TileInfo decode_tile(int tile, int m_blocks, int n_blocks, int total_kb, int split_k, int n_fast = 0) {
  TileInfo t;
  int rest;
  if (n_fast >= 2) {
    const int band = n_fast - 2;  // >= 1
    const int per_band = band * n_blocks;
    const int b = tile / per_band;
    const int in_band = tile - b * per_band;
    const int m0 = b * band;
    const int mb = min(band, m_blocks - m0);
    t.m_blk = m0 + in_band % mb;
    t.n_blk = in_band / mb;
    t.kb0 = 0;
    t.kb1 = total_kb;
    return t;
  }
  if (n_fast) {
    t.n_blk = tile % n_blocks;
    rest = tile / n_blocks;
    t.m_blk = rest % m_blocks;
    rest /= m_blocks;
  } else {
    t.m_blk = tile % m_blocks;
    rest = tile / m_blocks;
    t.n_blk = rest % n_blocks;
    rest /= n_blocks;
  }
  int split = rest;
  int kb_per = (total_kb + split_k - 1) / split_k;
  t.kb0 = split * kb_per;
  t.kb1 = min(total_kb, t.kb0 + kb_per);
  return t;
}

PG_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }


// Epilogue of one swap-AB tile for the non-GEGLU modes: thread = weight row (feature) fr, columns = tokens.  The mode is a
// template parameter and the output pointer is advanced by a precomputed byte stride, so that one element costs a
// handful of instructions (the epilogue of the decode GEMMs is issue-bound: 4 warps x 64 elements per tile).
template <int BN, int MODE>
PG_DEVINL void swap_tile_epilogue(const GemmArgs& args, uint32_t taddr, int fr, int j_base, bool first_split,
                                  int col_begin = 0, int col_end = BN) {
  const bool f_ok = fr < args.features;
  const int nvalid = min(min(BN, col_end), args.tokens - j_base);
  const float scale = args.scale;
  const float bias_s = (args.bias != nullptr && f_ok && first_split) ? __ldg(args.bias + fr) * scale : 0.f;
  constexpr int ESZ = MODE == PG_EPI_BF16 ? 2 : 4;
  char* dst = reinterpret_cast<char*>(args.out) + (static_cast<long long>(j_base + col_begin) * args.ldo + fr) * ESZ;
  const long long step = args.ldo * ESZ;
  const char* res = (MODE == PG_EPI_F32 && args.resid != nullptr)
                        ? reinterpret_cast<const char*>(args.resid + static_cast<long long>(j_base + col_begin) * args.ldr + fr) : nullptr;
  const long long rstep = args.ldr * 4;
  const bool gelu = MODE == PG_EPI_BF16 && args.act_gelu != 0;
#pragma unroll 1
  for (int c0 = col_begin; c0 < BN; c0 += 16) {
    if (c0 >= nvalid) break;  // warp-uniform
    uint32_t r[16];
    tmem_ld16(taddr + c0, r);
    tmem_ld_wait();
    if (!f_ok) continue;
    const int n = min(16, nvalid - c0);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < n) {
        float x = fmaf(__uint_as_float(r[i]), scale, bias_s);
        if (MODE == PG_EPI_BF16) {
          if (gelu) x = gelu_tanh_fast(x);
          *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16(x);
        } else if (MODE == PG_EPI_F32) {
          if (res != nullptr) x += *reinterpret_cast<const float*>(res);
          *reinterpret_cast<float*>(dst) = x;
        } else {
          atomicAdd(reinterpret_cast<float*>(dst), x);
        }
      }
      dst += step;
      if (MODE == PG_EPI_F32) res += rstep;
    }
  }
}

// lm_head epilogue (swap-AB, fp32 logits + bias) that also emits, per token column, the softmax statistics of the 32
// vocabulary rows this warp holds: (m, s) = (max_r x_r, sum_r exp2((x_r - m) * c)), c = inv_temperature * log2(e).  The
// sampler (sampler.cu) turns them into the row maximum, the partition function and the segment masses of its inverse-CDF
// draw without touching the 1 MB logit row: temperature scaling + the first two passes of softmax/top-p
// (inference.py:63-66,90-106) folded into the GEMM that produces the logits (modeling_gemma.py:523-525).
// Both warp reductions are single redux.sync instructions: the maximum on an order-preserving integer key, the sum on
// 2^-24 fixed point (every term is in [0, 1] and the maximum contributes exactly 2^24, so the sum of 32 terms is exact to
// 32 * 2^-25 relative) -- five-step shuffle butterflies for 64 columns x 2 quantities per tile made the epilogue, not the
// weight stream, pace the kernel (measured: 176 -> 195 us).  Rows past `features` (vocabulary tail) count as -inf; all 32
// lanes take part.  Layout: stats[segment][token] (one coalesced 8-byte-per-lane store per 16 columns).
PG_DEVINL uint32_t ordered_key(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
PG_DEVINL float ordered_key_inv(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); }

template <int BN>
PG_DEVINL void swap_tile_epilogue_f32_stats(const GemmArgs& args, uint32_t taddr, int fr, int j_base, int seg, int lane) {
  const bool f_ok = fr < args.features;
  const int nvalid = min(BN, args.tokens - j_base);
  const float scale = args.scale, c = args.stat_c;
  const float bias_s = (args.bias != nullptr && f_ok) ? __ldg(args.bias + fr) * scale : 0.f;
  float* dst = reinterpret_cast<float*>(args.out) + static_cast<long long>(j_base) * args.ldo + fr;
  float2* srow = args.stats + static_cast<long long>(seg) * args.stats_ld + j_base;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 16) {
    if (c0 >= nvalid) break;  // warp-uniform
    uint32_t r[16];
    tmem_ld16(taddr + c0, r);
    tmem_ld_wait();
    const int n = min(16, nvalid - c0);
    float keep_m = -INFINITY, keep_s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float x = f_ok ? fmaf(__uint_as_float(r[i]), scale, bias_s) : -INFINITY;
      if (i < n && f_ok) dst[static_cast<long long>(c0 + i) * args.ldo] = x;
      const float m = ordered_key_inv(__reduce_max_sync(0xffffffffu, ordered_key(x)));
      const float e = x == -INFINITY ? 0.f : exp2f((x - m) * c);
      const uint32_t sum = __reduce_add_sync(0xffffffffu, __float2uint_rn(e * 16777216.f));
      if (lane == i) { keep_m = m; keep_s = static_cast<float>(sum) * (1.0f / 16777216.f); }
    }
    if (lane < n) srow[c0 + lane] = make_float2(keep_m, keep_s);
  }
}

// Stores 32 rows x 128 bytes held one row per lane (pk = the lane's 8 16-byte pieces) as full 128-byte lines: the rows
// go through a warp-private swizzled 4 KB shared-memory tile and leave with 8 lanes per row, so one store instruction
// writes 4 complete lines instead of 32 16-byte fragments (the L2 write-request rate of the row-per-thread layout,
// 4096 requests per 128x256 bf16 tile, is what bounds the token-major epilogues: 923 vs 1490 TFLOP/s without stores).
PG_DEVINL void warp_store_rows_128B(uint32_t stage, int lane, const uint32_t (&pk)[32], char* gbase, long long pitch_bytes,
                                    int rows_valid, bool cols_ok) {
  __syncwarp();  // the previous user of the tile has read it out
  const uint32_t my_row = stage + lane * 128;
#pragma unroll
  for (int pc = 0; pc < 8; ++pc)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + ((pc ^ (lane & 7)) << 4)), "r"(pk[4 * pc]),
                 "r"(pk[4 * pc + 1]), "r"(pk[4 * pc + 2]), "r"(pk[4 * pc + 3]) : "memory");
  __syncwarp();
  const int sub = lane >> 3, piece = lane & 7;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r8 = 4 * k + sub;
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(stage + r8 * 128 + ((piece ^ (r8 & 7)) << 4)) : "memory");
    if (cols_ok && r8 < rows_valid) *reinterpret_cast<uint4*>(gbase + r8 * pitch_bytes + piece * 16) = v;
  }
}

// As warp_store_rows_128B, but every row has its own destination (pages of the KV cache): `my_row` = this lane's row base
// (null = skip the row).  The row pointers travel with the transposed lane mapping by shuffle.
PG_DEVINL void warp_store_rows_128B_ptr(uint32_t stage, int lane, const uint32_t (&pk)[32], char* my_row) {
  __syncwarp();
  const uint32_t row_s = stage + lane * 128;
#pragma unroll
  for (int pc = 0; pc < 8; ++pc)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_s + ((pc ^ (lane & 7)) << 4)), "r"(pk[4 * pc]),
                 "r"(pk[4 * pc + 1]), "r"(pk[4 * pc + 2]), "r"(pk[4 * pc + 3]) : "memory");
  __syncwarp();
  const int sub = lane >> 3, piece = lane & 7;
  const unsigned long long mine = reinterpret_cast<unsigned long long>(my_row);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r8 = 4 * k + sub;
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(stage + r8 * 128 + ((piece ^ (r8 & 7)) << 4)) : "memory");
    char* dst = reinterpret_cast<char*>(__shfl_sync(0xffffffffu, mine, r8));
    if (dst != nullptr) *reinterpret_cast<uint4*>(dst + piece * 16) = v;
  }
}

// sin / cos of a non-negative fp32 angle of up to a few thousand radians (position id x inverse frequency): three-constant
// Cody-Waite reduction by 2 pi (every step one FMA), then the hardware approximations on [-pi, pi] (abs. error ~5e-7).  The
// rotated values are rounded to bf16 (2^-9 relative) right away.
PG_DEVINL void sincos_2pi(float a, float& sn, float& cs) {
  const float n = rintf(a * 0.15915494309189535f);
  float r = fmaf(-n, 6.28125f, a);
  r = fmaf(-n, 1.935005187988281e-3f, r);
  r = fmaf(-n, 3.01991598195675e-7f, r);
  sn = __sinf(r);
  cs = __cosf(r);
}

// Epilogue of the prefill q/k/v projection (ROPE kernels): the N tile is exactly one head (BN == head_dim).  q and k heads
// are rotated (rotate-half pairs (i, i + BN/2), angle = pos[token] * inv_freq[i]: modeling_gemma.py:116-151) in registers,
// every head is written as bf16 into the dense [tokens, (Hq + 2 Hkv) * dh] buffer the prefill attention reads through
// strided tensor maps, and k / v heads are ALSO appended to their pages of the KV cache (KVCache.update, :18-57) -- the
// standalone RoPE + append kernel and its bf16 round trip of the projections are gone.  All stores are full 128-byte lines.
template <int BN>
PG_DEVINL void qkv_rope_tile_epilogue(const GemmArgs& args, uint32_t taddr, int tok, int head, uint32_t stage, int rows_valid,
                                      int lane, int warp_row0) {
  constexpr int HALF = BN / 2;
  const bool row_ok = tok < args.tokens;
  const bool is_q = head < args.rope_hq, is_v = head >= args.rope_hq + args.rope_hkv;
  __nv_bfloat16* out_bf = reinterpret_cast<__nv_bfloat16*>(args.out);
  char* dense0 = reinterpret_cast<char*>(out_bf + static_cast<long long>(warp_row0) * args.ldo + static_cast<long long>(head) * BN);
  char* page_row = nullptr;  // this token's row of the k / v head inside its cache page
  if (!is_q && args.k_pages != nullptr && row_ok) {
    const int b = tok / args.tokens_per_seq;
    const int slot = __ldg(args.slot_base + b) + (tok - b * args.tokens_per_seq);
    const long long page = __ldg(args.page_table + static_cast<long long>(b) * args.max_pages + (slot >> 6));
    const int hk = is_v ? head - args.rope_hq - args.rope_hkv : head - args.rope_hq;
    __nv_bfloat16* base = is_v ? args.v_pages : args.k_pages;
    page_row = reinterpret_cast<char*>(base + ((page * 64 + (slot & 63)) * args.rope_hkv + hk) * BN);
  }
  const float posf = row_ok ? static_cast<float>(__ldg(args.rope_pos + tok)) : 0.f;
  auto emit = [&](const uint32_t (&pk)[32], int col) {  // 64 columns starting at `col` of the head
    warp_store_rows_128B(stage, lane, pk, dense0 + col * 2, args.ldo * 2, rows_valid, true);
    if (!is_q && args.k_pages != nullptr) warp_store_rows_128B_ptr(stage, lane, pk, page_row != nullptr ? page_row + col * 2 : nullptr);
  };
  if (is_v) {
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 64) {
      uint32_t pk[32];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t r[16];
        tmem_ld16(taddr + c0 + 16 * j, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[8 * j + i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
      }
      emit(pk, c0);
    }
    return;
  }
  if constexpr (BN == 64) {
    // the head's 64 columns are one output row: [y1 (32) | y2 (32)]
    uint32_t pk[32];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t x1[16], x2[16];
      tmem_ld16(taddr + 16 * j, x1);
      tmem_ld16(taddr + HALF + 16 * j, x2);
      tmem_ld_wait();
      float y1[16], y2[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float sn, cs;
        sincos_2pi(posf * __ldg(args.rope_inv_freq + 16 * j + i), sn, cs);
        const float a = __uint_as_float(x1[i]), b = __uint_as_float(x2[i]);
        y1[i] = a * cs - b * sn;
        y2[i] = b * cs + a * sn;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        pk[8 * j + i] = pack_bf16(y1[2 * i], y1[2 * i + 1]);
        pk[16 + 8 * j + i] = pack_bf16(y2[2 * i], y2[2 * i + 1]);
      }
    }
    emit(pk, 0);
  } else {
#pragma unroll 1
    for (int g = 0; g < HALF / 64; ++g) {  // columns [64 g, 64 g + 64) pair with [HALF + 64 g, ...)
      uint32_t pk1[32], pk2[32];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t x1[16], x2[16];
        tmem_ld16(taddr + 64 * g + 16 * j, x1);
        tmem_ld16(taddr + HALF + 64 * g + 16 * j, x2);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float ya[2], yb[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float sn, cs;
            sincos_2pi(posf * __ldg(args.rope_inv_freq + 64 * g + 16 * j + 2 * i + e), sn, cs);
            const float a = __uint_as_float(x1[2 * i + e]), b = __uint_as_float(x2[2 * i + e]);
            ya[e] = a * cs - b * sn;
            yb[e] = b * cs + a * sn;
          }
          pk1[8 * j + i] = pack_bf16(ya[0], ya[1]);
          pk2[8 * j + i] = pack_bf16(yb[0], yb[1]);
        }
      }
      emit(pk1, 64 * g);
      emit(pk2, HALF + 64 * g);
    }
  }
}

// Epilogue of one token-major (non-swap) tile for the non-GEGLU modes: thread = token row, columns = features.
// 64 columns per step: the fp32 residual of the whole step is requested first (16 independent 16-byte loads per thread),
// then the accumulator is pulled out of TMEM, so that one global round trip covers 64 columns instead of 16 (the
// residual read is what bounds the short-K GEMMs: out_proj / o_proj).
template <int BN, int MODE>
PG_DEVINL void rowmajor_tile_epilogue(const GemmArgs& args, uint32_t taddr, int tok, int n0, bool first_split,
                                       uint32_t stage, int rows_valid, int lane, int c_begin = 0, int c_end = BN) {
  constexpr int STEP = BN >= 64 ? 64 : BN;
  constexpr int NCH = STEP / 16;
  const bool row_ok = tok < args.tokens;
  const float scale = args.scale;
  const bool has_bias = args.bias != nullptr && first_split;
  const bool gelu = MODE == PG_EPI_BF16 && args.act_gelu != 0;
  __nv_bfloat16* out_bf = reinterpret_cast<__nv_bfloat16*>(args.out) + static_cast<long long>(tok) * args.ldo;
  const int orow = (MODE == PG_EPI_F32 && args.out_row_map != nullptr && tok < args.tokens) ? __ldg(args.out_row_map + tok) : tok;
  float* out_f = reinterpret_cast<float*>(args.out) + static_cast<long long>(orow) * args.ldo;
  const float* res = (MODE == PG_EPI_F32 && args.resid != nullptr)
                         ? args.resid + static_cast<long long>(args.resid_mod > 0 ? tok % args.resid_mod : tok) * args.ldr : nullptr;
  // 16-byte vector access is possible when the row bases are 16 B aligned (n0 and the step are multiples of 16 columns)
  const bool vec_ok = MODE == PG_EPI_BF16 ? ((reinterpret_cast<uintptr_t>(out_bf) & 15) == 0)
                                          : (((reinterpret_cast<uintptr_t>(out_f) & 15) == 0) &&
                                             (res == nullptr || (reinterpret_cast<uintptr_t>(res) & 15) == 0));
  const bool bias_vec = has_bias && ((reinterpret_cast<uintptr_t>(args.bias) & 15) == 0);
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += STEP) {
    const int f0 = n0 + c0;
    if (f0 >= args.features) break;  // warp-uniform
    const bool full = (f0 + STEP <= args.features) && vec_ok;
    float4 rr[STEP / 4];
    if (MODE == PG_EPI_F32 && res != nullptr && full && row_ok) {
#pragma unroll
      for (int i = 0; i < STEP / 4; ++i) rr[i] = *reinterpret_cast<const float4*>(res + f0 + 4 * i);
    }
    uint32_t r[NCH][16];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) tmem_ld16(taddr + c0 + 16 * ch, r[ch]);
    tmem_ld_wait();
    // bf16, 64 full columns, aligned rows: warp-collective coalesced store (every lane takes part, valid row or not)
    const bool collective = MODE == PG_EPI_BF16 && STEP == 64 && __all_sync(0xffffffffu, (f0 + STEP <= args.features) && vec_ok);
    if (!row_ok && !collective) continue;
    uint32_t pkc[32];
    if (full || collective) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[ch][i]) * scale;
        if (has_bias) {
          if (bias_vec) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(args.bias + f0 + 16 * ch) + i);
              v[4 * i] = fmaf(b4.x, scale, v[4 * i]); v[4 * i + 1] = fmaf(b4.y, scale, v[4 * i + 1]);
              v[4 * i + 2] = fmaf(b4.z, scale, v[4 * i + 2]); v[4 * i + 3] = fmaf(b4.w, scale, v[4 * i + 3]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(__ldg(args.bias + f0 + 16 * ch + i), scale, v[i]);
          }
        }
        if (MODE == PG_EPI_BF16) {
          if (gelu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = gelu_tanh_fast(v[i]);
          }
          if (collective) {
#pragma unroll
            for (int i = 0; i < 8; ++i) pkc[(8 * ch + i) & 31] = pack_bf16(v[2 * i], v[2 * i + 1]);
          } else {
            uint4* d4 = reinterpret_cast<uint4*>(out_bf + f0 + 16 * ch);
            d4[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            d4[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
          }
        } else if (MODE == PG_EPI_F32) {
          float4* d4 = reinterpret_cast<float4*>(out_f + f0 + 16 * ch);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 o4 = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            if (res != nullptr) {
              const float4 q4 = rr[4 * ch + i];
              o4.x += q4.x; o4.y += q4.y; o4.z += q4.z; o4.w += q4.w;
            }
            d4[i] = o4;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(out_f + f0 + 16 * ch + i, v[i]);
        }
      }
      if (collective)
        warp_store_rows_128B(stage, lane, pkc, reinterpret_cast<char*>(out_bf - static_cast<long long>(lane) * args.ldo + f0),
                             args.ldo * 2, rows_valid, true);
    } else {  // feature tail / unaligned rows: element by element
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int f = f0 + 16 * ch + i;
          if (f < args.features) {
            float x = __uint_as_float(r[ch][i]) * scale;
            if (has_bias) x = fmaf(__ldg(args.bias + f), scale, x);
            if (MODE == PG_EPI_BF16) {
              if (gelu) x = gelu_tanh_fast(x);
              out_bf[f] = __float2bfloat16(x);
            } else if (MODE == PG_EPI_F32) {
              out_f[f] = x + (res != nullptr ? res[f] : 0.f);
            } else {
              atomicAdd(out_f + f, x);
            }
          }
        }
      }
    }
  }
}

// fp32 (+bias, +fp32 residual) epilogue of a token-major tile with COALESCED global access.  The accumulator comes out
// of TMEM one row per thread; 32 columns at a time it is transposed through a warp-private, swizzled 4 KB shared-memory
// tile into the mapping "8 lanes per row, 16 bytes per lane", so that every global load / store instruction of the
// residual stream touches 4 full 128-byte lines instead of 32 partial ones (the row-per-thread version is bound by the
// LSU's line throughput: 16 k cycles per 128x256 tile, more than the main loop of the K <= 2048 GEMMs).
// The residual of slice s+1 is requested before slice s is finished.  Needs features % 4 == 0 and 16 B aligned rows.
template <int BN>
PG_DEVINL void rowmajor_tile_epilogue_f32_coalesced(const GemmArgs& args, uint32_t taddr, int m0, int n0, bool first_split,
                                                    uint32_t stage, int q, int lane, int sl_begin = 0, int sl_end = BN / 32) {
  const float scale = args.scale;
  const bool has_bias = args.bias != nullptr && first_split;
  float* out = reinterpret_cast<float*>(args.out);
  const float* res = args.resid;
  const int row_base = m0 + q * 32;
  const int sub = lane >> 3, piece = lane & 7;
  const int nsl = min(sl_end, (min(BN, args.features - n0) + 31) / 32);
  auto load_res = [&](int sl, float4 (&dst)[8]) {
    const int col = n0 + sl * 32 + piece * 4;
    const bool col_ok = col < args.features;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int row = row_base + 4 * k + sub;
      const int rrow = args.resid_mod > 0 ? row % args.resid_mod : row;
      dst[k] = (col_ok && row < args.tokens) ? *reinterpret_cast<const float4*>(res + static_cast<long long>(rrow) * args.ldr + col)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float4 rr[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) rr[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (res != nullptr && sl_begin < nsl) load_res(sl_begin, rr);
#pragma unroll 1
  for (int sl = sl_begin; sl < nsl; ++sl) {
    uint32_t a0[16], a1[16];
    tmem_ld16(taddr + sl * 32, a0);
    tmem_ld16(taddr + sl * 32 + 16, a1);
    tmem_ld_wait();
    __syncwarp();  // the previous slice has been read out of the staging tile
    const uint32_t my_row = stage + lane * 128;
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + ((pc ^ (lane & 7)) << 4)), "r"(a0[4 * pc]),
                   "r"(a0[4 * pc + 1]), "r"(a0[4 * pc + 2]), "r"(a0[4 * pc + 3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + (((pc + 4) ^ (lane & 7)) << 4)), "r"(a1[4 * pc]),
                   "r"(a1[4 * pc + 1]), "r"(a1[4 * pc + 2]), "r"(a1[4 * pc + 3]) : "memory");
    }
    __syncwarp();
    float4 rn[8];
    if (res != nullptr && sl + 1 < nsl) load_res(sl + 1, rn);
    const int col = n0 + sl * 32 + piece * 4;
    const bool col_ok = col < args.features;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_bias && col_ok) {
      b4 = __ldg(reinterpret_cast<const float4*>(args.bias + col));
      b4.x *= scale; b4.y *= scale; b4.z *= scale; b4.w *= scale;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r8 = 4 * k + sub;
      const int row = row_base + r8;
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "r"(stage + r8 * 128 + ((piece ^ (r8 & 7)) << 4)) : "memory");
      if (col_ok && row < args.tokens) {
        float4 o4;
        o4.x = fmaf(v.x, scale, b4.x) + rr[k].x; o4.y = fmaf(v.y, scale, b4.y) + rr[k].y;
        o4.z = fmaf(v.z, scale, b4.z) + rr[k].z; o4.w = fmaf(v.w, scale, b4.w) + rr[k].w;
        const int orow = args.out_row_map != nullptr ? __ldg(args.out_row_map + row) : row;
        *reinterpret_cast<float4*>(out + static_cast<long long>(orow) * args.ldo + col) = o4;
      }
    }
    if (res != nullptr && sl + 1 < nsl) {
#pragma unroll
      for (int k = 0; k < 8; ++k) rr[k] = rn[k];
    }
  }
}

// GeGLU epilogue of a swap-AB tile for ONE group of four epilogue warps that owns CG token columns starting at column
// `cb` of the tile (CG = BN with four epilogue warps, BN / 2 with eight).  Packed weight rows: [64 gate | 64 up] per
// 128-row tile, so warps q<2 hold the gate rows and warps q>=2 the matching up rows of the same 64 output features.
// The two halves of the group's columns are exchanged through shared memory (gate warps finish [0, CG/2), up warps
// [CG/2, CG)), the bf16 results are transposed through shared memory and leave as 16-byte vectors (one 128-byte row of
// 64 features per token).  `xg` = the group's exchange area (64 * CG floats), `bar` = its named barrier.
template <int CG>
PG_DEVINL void geglu_swap_epilogue_group(const GemmArgs& args, uint32_t taddr, float* xg, int bar, int q, int lane, int et,
                                         int j0, int m_blk) {
  constexpr int HB = CG / 2;
  constexpr int LDN = HB >= 16 ? 16 : 8;
  __nv_bfloat16* out_bf = reinterpret_cast<__nv_bfloat16*>(args.out);
  const bool is_gate = q < 2;
  const int r = (q * 32 + lane) & 63;
  float* xsend = xg + (is_gate ? 64 * HB : 0);  // gate warps fill the second half of the area, up warps the first
  float* xrecv = xg + (is_gate ? 0 : 64 * HB);
  const int send0 = is_gate ? HB : 0, keep0 = is_gate ? 0 : HB;
#pragma unroll 1
  for (int c0 = 0; c0 < HB; c0 += LDN) {
    if (j0 + send0 + c0 >= args.tokens) break;
    uint32_t v[16];
    if constexpr (LDN == 16) tmem_ld16(taddr + send0 + c0, v); else tmem_ld8(taddr + send0 + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < LDN; ++i) xsend[r * HB + ((c0 + i) ^ (r & (HB - 1) & 31))] = __uint_as_float(v[i]);
  }
  named_bar_sync(bar, 128);
  uint32_t pk[HB / 2];
#pragma unroll
  for (int c0 = 0; c0 < HB; c0 += LDN) {
    uint32_t v[16];
    if (j0 + keep0 + c0 < args.tokens) {
      if constexpr (LDN == 16) tmem_ld16(taddr + keep0 + c0, v); else tmem_ld8(taddr + keep0 + c0, v);
      tmem_ld_wait();
    }
#pragma unroll
    for (int i = 0; i < LDN; i += 2) {
      float res[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float mine = __uint_as_float(v[i + e]);
        float other = xrecv[r * HB + ((c0 + i + e) ^ (r & (HB - 1) & 31))];
        res[e] = is_gate ? gelu_tanh_fast(mine) * other : gelu_tanh_fast(other) * mine;
      }
      pk[(c0 + i) / 2] = pack_bf16(res[0], res[1]);
    }
  }
  named_bar_sync(bar, 128);  // every exchange value has been consumed: the area is reused for the transposed tile
  __nv_bfloat16* out_s = reinterpret_cast<__nv_bfloat16*>(xg);  // [CG tokens][64 features]
#pragma unroll
  for (int c = 0; c < HB; c += 2) {
    const __nv_bfloat162 two = *reinterpret_cast<const __nv_bfloat162*>(&pk[c / 2]);
    out_s[(keep0 + c) * 64 + r] = two.x;
    out_s[(keep0 + c + 1) * 64 + r] = two.y;
  }
  named_bar_sync(bar, 128);
  {
    const int f0 = m_blk * 64;
    if (f0 < args.features / 2) {
#pragma unroll 1
      for (int pc = et; pc < CG * 8; pc += 128) {
        const int j = j0 + (pc >> 3), part = pc & 7;
        if (j < args.tokens)
          *reinterpret_cast<uint4*>(out_bf + static_cast<long long>(j) * args.ldo + f0 + part * 8) =
              *reinterpret_cast<const uint4*>(out_s + (pc >> 3) * 64 + part * 8);
      }
    }
  }
  named_bar_sync(bar, 128);
}
template <int BN>
PG_DEVINL void geglu_swap_epilogue(const GemmArgs& args, uint32_t taddr, float* xch, int bar, int q, int lane, int et,
                                   int j_base, int m_blk) {
  geglu_swap_epilogue_group<BN>(args, taddr, xch, bar, q, lane, et, j_base, m_blk);
}

// Epilogue of one token-major 128 x BN accumulator tile (TMEM lanes = tokens): the per-call-site variants of the prefill GEMMs.
// Shared by the one-CTA kernel and the CTA-pair kernel (whose two CTAs each own 128 of the pair's 256 token rows).
// EPI8: eight epilogue warps, two per TMEM lane quadrant, each takes half of the tile's columns.
template <int BN, bool EPI8, bool ROPE>
PG_DEVINL void token_major_epilogue(const GemmArgs& args, uint32_t taddr, int m_blk, int n_blk, bool first_split, int warp, int lane,
                                    uint32_t xch_base) {
  const int q = warp & 3;
  const int rl = q * 32 + lane;
  const int mode = args.mode;
  __nv_bfloat16* out_bf = reinterpret_cast<__nv_bfloat16*>(args.out);
  const int tok = m_blk * BM + rl;
  const bool row_ok = tok < args.tokens;
  // eight epilogue warps: two per TMEM lane quadrant, each takes half of the tile's columns
  const int half = EPI8 ? ((warp - 2) >> 2) : 0;
  const uint32_t wstage = xch_base + (warp - 2) * 4096;  // this warp's transposition tile
  const int rows_valid = max(0, min(32, args.tokens - (m_blk * BM + q * 32)));
  // column range of this warp: halves of >= 64 columns (BN = 64: the first warp of the pair takes everything)
  const int cb = !EPI8 ? 0 : (BN >= 128 ? half * (BN / 2) : (half == 0 ? 0 : BN));
  const int ce = !EPI8 ? BN : (BN >= 128 ? cb + BN / 2 : BN);
  if constexpr (ROPE) {
    qkv_rope_tile_epilogue<BN>(args, taddr, tok, n_blk, wstage, rows_valid, lane, m_blk * BM + q * 32);
  } else if (mode == PG_EPI_GEGLU) {
    // columns: [g0..g63 | u0..u63] per 128-column block; out feature = n_blk*BN/2 + blk*64 + c.  The 64 bf16
    // results of a block (one 128-byte row per token) leave through the warp transposition tile as full lines.
    const bool al = ((reinterpret_cast<uintptr_t>(args.out) & 15) == 0) && ((args.ldo % 8) == 0);
#pragma unroll 1
    for (int blk = ((EPI8 && BN >= 256) ? half : 0); blk < ((EPI8 && BN >= 256) ? half + 1 : (half == 0 ? BN / 128 : 0)); ++blk) {
      const int f0 = n_blk * (BN / 2) + blk * 64;
      if (f0 >= args.features / 2) break;  // warp-uniform (features % 128 == 0)
      uint32_t pk[32];
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t g[16], u[16];
        tmem_ld16(taddr + blk * 128 + c0, g);
        tmem_ld16(taddr + blk * 128 + 64 + c0, u);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a = gelu_tanh_fast(__uint_as_float(g[2 * i])) * __uint_as_float(u[2 * i]);
          float b = gelu_tanh_fast(__uint_as_float(g[2 * i + 1])) * __uint_as_float(u[2 * i + 1]);
          pk[(c0 / 2 + i) & 31] = pack_bf16(a, b);
        }
      }
      if (al) {
        warp_store_rows_128B(wstage, lane, pk, reinterpret_cast<char*>(out_bf + static_cast<long long>(m_blk * BM + q * 32) * args.ldo + f0),
                             args.ldo * 2, rows_valid, true);
      } else if (row_ok) {
        __nv_bfloat16* dst = out_bf + static_cast<long long>(tok) * args.ldo + f0;
#pragma unroll
        for (int i = 0; i < 32; ++i) *reinterpret_cast<uint32_t*>(dst + 2 * i) = pk[i];
      }
    }
  } else if (mode == PG_EPI_BF16) {
    rowmajor_tile_epilogue<BN, PG_EPI_BF16>(args, taddr, tok, n_blk * BN, first_split, wstage, rows_valid, lane, cb, ce);
  } else if (mode == PG_EPI_F32) {
    // (measured: pulling the NEXT tile's residual rows into L2 from here does not help: o_proj 926 -> 824 TFLOP/s)
    if (args.f32_coalesced)
      rowmajor_tile_epilogue_f32_coalesced<BN>(args, taddr, m_blk * BM, n_blk * BN, first_split,
                                               wstage, q, lane, cb / 32, ce / 32);
    else
      rowmajor_tile_epilogue<BN, PG_EPI_F32>(args, taddr, tok, n_blk * BN, first_split, wstage, rows_valid, lane, cb, ce);
  } else {
    rowmajor_tile_epilogue<BN, PG_EPI_ATOMIC_F32>(args, taddr, tok, n_blk * BN, first_split, wstage, rows_valid, lane, cb, ce);
  }
}

// SPLITK = true: swap-AB kernel specialised for the split-K red.add epilogue with EIGHT epilogue warps (two per TMEM lane
// quadrant, half of the token columns each).  With one CTA per SM (qkv / o_proj: ~144 CTAs) the 64 dependent red
// instructions per thread are the serial tail of the launch; two warps per quadrant halve it.  Everything but the
// red.add epilogue is compiled out, which keeps the 320-thread CTA at two per SM.
template <int BN, bool SWAP, bool SPLITK = false, bool ROPE = false>
__global__ void __launch_bounds__(SPLITK ? 320 : NUM_THREADS, two_per_sm(BN, SWAP) ? 2 : 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                    const GemmArgs args) {
  constexpr int NEPI = SPLITK ? 8 : 4;  // epilogue warps
  // (token-major kernel: SPLITK = true selects the same 320-thread shape, eight epilogue warps that split the columns;
  //  it pays for the short-K SigLIP GEMMs whose epilogue outlasts the main loop, and costs registers on the long-K ones)
  constexpr int STAGES = num_stages(BN, SWAP);
  static_assert(STAGES >= 3, "pipeline too shallow");
  constexpr int STAGE_BYTES = stage_bytes(BN);
  constexpr int TMEM_COLS = tmem_cols(BN);
  constexpr uint32_t IDESC = make_idesc_bf16(BM, BN);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 128B swizzle atoms need 1024 B alignment
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  uint8_t* smem_gen = smem_raw;
  float* xch = reinterpret_cast<float*>(smem_gen + STAGES * STAGE_BYTES);
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES + xch_bytes(BN, SWAP);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + xch_bytes(BN, SWAP) + 8 * (2 * STAGES + 2 * ACC_STAGES));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int rowsA = SWAP ? args.features : args.tokens;
  const int rowsB = SWAP ? args.tokens : args.features;
  const int m_blocks = (rowsA + BM - 1) / BM;
  const int n_blocks = (rowsB + BN - 1) / BN;
  const int total_kb = (args.K + BK - 1) / BK;
  const int num_tiles = m_blocks * n_blocks * args.split_k;

  const bool tr = (args.trace != nullptr) && blockIdx.x == 0;
  if (tr && threadIdx.x == 0) args.trace[0] = clock64();  // kernel entry
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), SPLITK ? 8 : 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // =================================== TMA producer ===================================
    if (lane == 0) {
      // weights are streamed once in decode (evict first); activations are re-read by every CTA (evict last)
      const uint64_t hintA = SWAP ? kEvictFirst : kEvictNormal;
      const uint64_t hintB = (SWAP || args.n_fast == 1) ? kEvictLast : kEvictNormal;
      int stage = 0;
      uint32_t phase = 0;
      // The weight operand never depends on the previous kernel: with programmatic dependent launch its first
      // pipeline stages are fetched BEFORE griddepcontrol.wait, i.e. while the previous kernel is still running.
      // (token-major kernel: the weights are operand B; only the first tile of a CTA can be prefetched this way, which is
      //  what matters for the few-token latency path where a CTA has one tile and the launch chain is the critical path)
      int pre = 0;
      if (static_cast<int>(blockIdx.x) < num_tiles) {
        const TileInfo t = decode_tile(blockIdx.x, m_blocks, n_blocks, total_kb, args.split_k, args.n_fast);
        pre = min(STAGES, t.kb1 - t.kb0);
        for (int s = 0; s < pre; ++s) {
          mbar_expect_tx(full_bar(s), STAGE_BYTES);
          if (SWAP) tma_load_2d(smem_base + s * STAGE_BYTES, &tmapA, full_bar(s), (t.kb0 + s) * BK, t.m_blk * BM, hintA);
          else tma_load_2d(smem_base + s * STAGE_BYTES + A_TILE_BYTES, &tmapB, full_bar(s), (t.kb0 + s) * BK, t.n_blk * BN, hintB);
        }
      }
      if (tr) args.trace[1] = clock64();  // weights prefetch issued
      // (triggering before the wait, as the decode RMSNorm does, changes nothing here: measured for q/k/v, o, gate||up, down)
      griddep_wait();
      if (tr) args.trace[2] = clock64();  // previous kernel complete
      griddep_launch_dependents();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileInfo t = decode_tile(tile, m_blocks, n_blocks, total_kb, args.split_k, args.n_fast);
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const bool weights_in_flight = pre > 0;  // this stage's weight tile was requested before the dependency wait
          if (weights_in_flight) {
            --pre;
          } else {
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          }
          if (!(SWAP && weights_in_flight)) tma_load_2d(sa, &tmapA, full_bar(stage), kb * BK, t.m_blk * BM, hintA);
          if (!(!SWAP && weights_in_flight)) tma_load_2d(sa + A_TILE_BYTES, &tmapB, full_bar(stage), kb * BK, t.n_blk * BN, hintB);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();  // the whole warp reaches the CTA-wide barrier below together (bar.sync counts warps, not lanes)
  } else if (warp == 1) {
    // =================================== MMA issuer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileInfo t = decode_tile(tile, m_blocks, n_blocks, total_kb, args.split_k, args.n_fast);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = make_sdesc_k_sw128(sa);
          const uint64_t bdesc = make_sdesc_k_sw128(sa + A_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128B swizzle row: +2 in the (addr>>4) field
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb > t.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs have read it
          if (kb == t.kb1 - 1) umma_commit(tfull_bar(acc));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =================================== epilogue warps =================================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int et = (warp - 2) * 32 + lane;
    griddep_wait();          // outputs / residual / bias / fp32 activations may depend on the previous kernel
    if constexpr (SWAP) {
      if (args.zero_buf != nullptr) {
        const long long n4 = args.zero_count >> 2;  // (host: zero_count % 4 == 0, 16-byte aligned)
        const long long per = (n4 + gridDim.x - 1) / gridDim.x;
        const long long lo = blockIdx.x * per, hi = min(n4, lo + per);
        for (long long i = lo + et; i < hi; i += NEPI * 32) reinterpret_cast<float4*>(args.zero_buf)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    const int mode = args.mode;
    __nv_bfloat16* out_bf = reinterpret_cast<__nv_bfloat16*>(args.out);
    float* out_f = reinterpret_cast<float*>(args.out);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileInfo t = decode_tile(tile, m_blocks, n_blocks, total_kb, args.split_k, args.n_fast);
      mbar_wait(tfull_bar(acc), acc_phase);
      if (tr && threadIdx.x == 64) args.trace[3] = clock64();  // accumulator ready (first tile)
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      const int rl = q * 32 + lane;  // row inside the tile
      const bool first_split = (t.kb0 == 0);

      if constexpr (SPLITK && SWAP) {
        // eight epilogue warps: warps 2-5 take the token columns [0, BN/2), warps 6-9 the rest (BN >= 32)
        constexpr int HALF_COLS = BN >= 32 ? BN / 2 : BN;
        const int half = (warp - 2) >> 2;
        if (mode == PG_EPI_GEGLU) {
          if constexpr (BN >= 32)
            geglu_swap_epilogue_group<HALF_COLS>(args, taddr + half * HALF_COLS, xch + half * 64 * HALF_COLS, 1 + half, q, lane,
                                                 ((warp - 2) & 3) * 32 + lane, t.n_blk * BN + half * HALF_COLS, t.m_blk);
        } else if (half * HALF_COLS < BN) {
          swap_tile_epilogue<BN, PG_EPI_ATOMIC_F32>(args, taddr, t.m_blk * BM + rl, t.n_blk * BN, first_split,
                                                    half * HALF_COLS, (half + 1) * HALF_COLS);
        }
      } else if constexpr (!SWAP) {
        token_major_epilogue<BN, SPLITK, ROPE>(args, taddr, t.m_blk, t.n_blk, first_split, warp, lane, smem_base + STAGES * STAGE_BYTES);
      } else {
        // SWAP: this thread owns weight row (feature) fr; columns are tokens
        const int fr = t.m_blk * BM + rl;
        const int j_base = t.n_blk * BN;
        if (mode == PG_EPI_GEGLU) {
          geglu_swap_epilogue<BN>(args, taddr, xch, 1, q, lane, (warp - 2) * 32 + lane, j_base, t.m_blk);
        } else {
          // (measured: 16-byte REDG.F32x4 after a lane-quad transpose is ~2x SLOWER here than 4-byte coalesced reds)
          if (mode == PG_EPI_ATOMIC_F32) swap_tile_epilogue<BN, PG_EPI_ATOMIC_F32>(args, taddr, fr, j_base, first_split);
          else if (mode == PG_EPI_F32 && args.stats != nullptr) swap_tile_epilogue_f32_stats<BN>(args, taddr, fr, j_base, t.m_blk * 4 + q, lane);
          else if (mode == PG_EPI_F32) swap_tile_epilogue<BN, PG_EPI_F32>(args, taddr, fr, j_base, first_split);
          else swap_tile_epilogue<BN, PG_EPI_BF16>(args, taddr, fr, j_base, first_split);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (tr && threadIdx.x == 64) args.trace[4] = clock64();  // epilogue of this tile issued
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tr && threadIdx.x == 0) args.trace[5] = clock64();  // all roles done
  if (tr && threadIdx.x == 64) args.trace[6] = clock64();
  if (tr && threadIdx.x == 32) args.trace[7] = clock64();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// CTA-pair kernel (prefill, token-major, 256 x 256 output tile per pair of SMs): tcgen05.mma.cta_group::2.
// Each CTA of a 2-CTA cluster stages its 128 token rows of A and HALF of the 256 weight rows of B per k-block (32 KB per
// stage instead of 48 KB: a third less L2 -> shared-memory traffic per FLOP, which is what the power-capped prefill pays
// for), the leader issues 256 x 256 x 16 UMMAs that read both CTAs' shared memory, and every CTA's tensor memory receives its
// own 128 x 256 fp32 accumulator, which its epilogue warps drain exactly like a tile of the one-CTA kernel.
//   full[s]   (leader)     : 1 arrival (leader's expect_tx of 64 KB) + the bytes of BOTH CTAs' TMA loads
//   empty[s]  (both CTAs)  : the leader's tcgen05.commit, multicast to the pair
//   tfull[a]  (both CTAs)  : the leader's tcgen05.commit after the last k-block of a tile, multicast
//   tempty[a] (leader)     : 2 x NEPI arrivals, the epilogue warps of both CTAs
// ------------------------------------------------------------------------------------------------------------
constexpr int PAIR_BN = 256;
constexpr int PAIR_STAGE_BYTES = A_TILE_BYTES + (PAIR_BN / 2) * BK * 2;  // 32 KB
constexpr int PAIR_STAGES = (232448 - 256 - 8 * 4096) / PAIR_STAGE_BYTES;  // 6
constexpr int PAIR_SMEM = PAIR_STAGES * PAIR_STAGE_BYTES + 8 * 4096 + 256;

template <bool EPI8, bool ROPE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EPI8 ? 320 : NUM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const GemmArgs args) {
  constexpr int BN = PAIR_BN, STAGES = PAIR_STAGES, STAGE_BYTES = PAIR_STAGE_BYTES;
  constexpr int NEPI = EPI8 ? 8 : 4;
  constexpr uint32_t IDESC = make_idesc_bf16(2 * BM, BN);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t xch_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = xch_base + 8 * 4096;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + STAGES * STAGE_BYTES + 8 * 4096 + 8 * (2 * STAGES + 2 * ACC_STAGES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const bool leader = rank == 0;
  const int m_blocks = (args.tokens + 2 * BM - 1) / (2 * BM);  // 256-token blocks
  const int n_blocks = (args.features + BN - 1) / BN;
  const int total_kb = (args.K + BK - 1) / BK;
  const int num_tiles = m_blocks * n_blocks;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * NEPI);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr_addr, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised before either signals the other's
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // =================================== TMA producer (both CTAs) ===================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // feature tiles first (n_fast = 1): the weight matrix is the L2-resident operand (evict last), the activations pass through
      // (evict-first on the activations is WRONG here: the CTAs of the feature tiles that share a token tile read it at slightly
      //  different times, and an evict-first line is gone before the second reader arrives -- measured 2.97 -> 6.99 GB)
      const uint64_t hintA = kEvictNormal;
      const uint64_t hintB = args.n_fast == 1 ? kEvictLast : kEvictNormal;
      griddep_wait();
      griddep_launch_dependents();
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const TileInfo t = decode_tile(tile, m_blocks, n_blocks, total_kb, 1, args.n_fast);
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
          tma_load_2d_pair(sa, &tmapA, full_bar(stage), kb * BK, (2 * t.m_blk + rank) * BM, hintA);
          tma_load_2d_pair(sa + A_TILE_BYTES, &tmapB, full_bar(stage), kb * BK, t.n_blk * BN + rank * (BN / 2), hintB);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =================================== MMA issuer (leader CTA only) ===============================
    if (leader && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = make_sdesc_k_sw128(sa);
          const uint64_t bdesc = make_sdesc_k_sw128(sa + A_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(empty_bar(stage));
          if (kb == total_kb - 1) umma_commit_pair(tfull_bar(acc));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =================================== epilogue warps (both CTAs) =================================
    const int q = warp & 3;
    griddep_wait();
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const TileInfo t = decode_tile(tile, m_blocks, n_blocks, total_kb, 1, args.n_fast);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      token_major_epilogue<BN, EPI8, ROPE>(args, taddr, 2 * t.m_blk + rank, t.n_blk, true, warp, lane, xch_base);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory and write its tensor memory until the very end
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

template <bool EPI8, bool ROPE>
static int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, int num_tiles, cudaStream_t st) {
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(gemm_pair_kernel<EPI8, ROPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM) != cudaSuccess) {
      cudaGetLastError();
      return PG_ERR_CUDA;
    }
    configured[dev] = true;
  }
  const int pairs = num_sms() / 2;
  const int grid = 2 * (num_tiles < pairs ? num_tiles : pairs);
  return launch_kernel(gemm_pair_kernel<EPI8, ROPE>, dim3(grid), dim3(EPI8 ? 320 : NUM_THREADS), PAIR_SMEM, st, ta, tb, a) == cudaSuccess
             ? PG_OK : PG_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static long long* g_gemm_trace = nullptr;
static int g_gemm_trace_idx = 0;
extern "C" int pg_debug_set_gemm_trace(long long* p) { g_gemm_trace = p; g_gemm_trace_idx = 0; return 0; }
static int g_epi8_mode = 0;  // tuning: 1 = eight epilogue warps for every pair GEMM, 2 = four, 0 = by K
static int g_band_mode = 1;        // 0: never use the banded raster (A/B runs)
static int g_pair_mode = 1;        // 0: never use the CTA-pair kernel (A/B runs)
static int g_pair_min_tiles = 74;  // at least one 256 x 256 tile per pair of SMs
extern "C" int pg_debug_set_gemm_pair(int mode, int min_tiles) {
  g_pair_mode = mode & 1;
  g_epi8_mode = (mode >> 2) & 3;
  g_band_mode = (mode & 2) ? 0 : 1;  // bit 1: banded raster off
  if (min_tiles > 0) g_pair_min_tiles = min_tiles;
  return 0;
}
static int g_force_bn = 0;  // tuning sweeps only: 64 / 128 / 256 overrides the token-major kernel's N tile, 0 = automatic
extern "C" int pg_debug_set_gemm_bn(int bn) {
  if (bn != 0 && bn != 64 && bn != 128 && bn != 256) return PG_ERR_ARG;
  g_force_bn = bn;
  return 0;
}

template <int BN, bool SWAP, bool SPLITK = false, bool ROPE = false>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, int num_tiles, cudaStream_t st) {
  static bool configured[kMaxDevices] = {};
  constexpr int smem = smem_bytes(BN, SWAP);
  const int dev = current_device();
  if (!configured[dev]) {  // function attributes are per device
    if (cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, SWAP, SPLITK, ROPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();  // do not leave a sticky error behind
      return PG_ERR_CUDA;
    }
    configured[dev] = true;
  }
  const int slots = num_sms() * (two_per_sm(BN, SWAP) ? 2 : 1);
  const int grid = num_tiles < slots ? num_tiles : slots;
  return launch_kernel(gemm_tcgen05_kernel<BN, SWAP, SPLITK, ROPE>, dim3(grid), dim3(SPLITK ? 320 : NUM_THREADS), smem, st, ta, tb, a) == cudaSuccess
             ? PG_OK : PG_ERR_CUDA;
}

}  // namespace pg

using namespace pg;

extern "C" int pg_gemm_bf16(const void* x, long long ldx, const void* w, long long ldw, void* out, long long ldo,
                            const float* bias, const float* resid, long long ldr, int tokens, int features, int K,
                            int mode, int act_gelu, float scale, int swap, int split_k, void* stream) {
  return pg_gemm_bf16_fused(x, ldx, w, ldw, out, ldo, bias, resid, ldr, tokens, features, K, mode, act_gelu, scale, swap,
                            split_k, nullptr, stream);
}

extern "C" int pg_gemm_bf16_fused(const void* x, long long ldx, const void* w, long long ldw, void* out, long long ldo,
                                  const float* bias, const float* resid, long long ldr, int tokens, int features, int K,
                                  int mode, int act_gelu, float scale, int swap, int split_k, const PgGemmFusion* fu,
                                  void* stream) {
  if (tokens <= 0 || features <= 0 || K <= 0) return PG_ERR_ARG;
  if ((K % 8) != 0 || (ldw % 8) != 0) return PG_ERR_ARG;  // TMA: 16 B pitch granularity
  if (reinterpret_cast<uintptr_t>(w) & 15) return PG_ERR_ARG;
  if (x == nullptr || (ldx % 8) != 0 || (reinterpret_cast<uintptr_t>(x) & 15)) return PG_ERR_ARG;
  if (mode < PG_EPI_BF16 || mode > PG_EPI_GEGLU) return PG_ERR_ARG;
  if (mode == PG_EPI_GEGLU && (features % 128) != 0) return PG_ERR_ARG;
  if (swap < 0) swap = tokens <= 128 ? 1 : 0;
  if (swap && tokens > 128) return PG_ERR_ARG;
  const int total_kb = (K + BK - 1) / BK;
  if (split_k <= 0) split_k = 1;
  if (mode != PG_EPI_ATOMIC_F32) split_k = 1;
  if (split_k > total_kb) split_k = total_kb;
  {  // no empty splits
    int kb_per = (total_kb + split_k - 1) / split_k;
    split_k = (total_kb + kb_per - 1) / kb_per;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  GemmArgs a = {};
  a.tokens = tokens; a.features = features; a.K = K; a.split_k = split_k; a.mode = mode; a.act_gelu = act_gelu;
  a.scale = scale; a.out = out; a.ldo = ldo; a.bias = bias; a.resid = resid; a.ldr = ldr;
  if (fu != nullptr && fu->resid_row_mod > 0) {
    if (swap || resid == nullptr || mode != PG_EPI_F32) return PG_ERR_ARG;  // a token-major fp32 epilogue feature
    a.resid_mod = fu->resid_row_mod;
  }
  if (fu != nullptr && fu->out_row_map != nullptr) {
    if (swap || resid != nullptr || mode != PG_EPI_F32) return PG_ERR_ARG;  // token-major fp32 epilogue, no in-place residual
    a.out_row_map = fu->out_row_map;
  }
  if (fu != nullptr && (fu->zero_count > 0 || fu->stats != nullptr)) {
    if (!swap) return PG_ERR_ARG;  // these belong to the decode (swap-AB) kernels
    if (fu->zero_count > 0) {
      if (fu->zero_buf == nullptr || (fu->zero_count % 4) != 0 || (reinterpret_cast<uintptr_t>(fu->zero_buf) & 15)) return PG_ERR_ARG;
      a.zero_buf = fu->zero_buf; a.zero_count = fu->zero_count;
    }
    if (fu->stats != nullptr) {
      // plain fp32 logits + bias only (no residual, no in-kernel norm factor), one CTA per output tile
      if (mode != PG_EPI_F32 || resid != nullptr || split_k != 1 || fu->stats_ld < tokens ||
          (reinterpret_cast<uintptr_t>(fu->stats) & 7) || !(fu->stat_c > 0.f))
        return PG_ERR_ARG;
      a.stats = static_cast<float2*>(fu->stats); a.stats_ld = fu->stats_ld; a.stat_c = fu->stat_c;
    }
  }
  a.n_fast = 0;
  double band_rows = 0;  // banded raster: token rows per band (converted to tiles of the kernel that is chosen below)
  if (!swap && split_k == 1) {  // DRAM traffic estimate of the raster orders
    // an operand stays L2 resident while the other one streams through only if it is well below the 126 MB: 48 MB budget
    // (measured: a 68 MB activation matrix under a 134 MB weight stream was re-read 7 times)
    const double l2 = 48e6, A = 2.0 * tokens * K, Bw = 2.0 * features * K;
    const int mb = (tokens + BM - 1) / BM;
    const int nb256 = (features + 255) / 256;
    const double m_fast = (A > l2 ? A * nb256 : A) + Bw;
    const double n_fast = A + (Bw > l2 ? Bw * ((static_cast<double>(mb) * nb256 + 147) / 148) : Bw);
    const double rows32 = 32e6 / (2.0 * K);  // token rows whose activations fill 32 MB
    const double banded = A + Bw * ceil(tokens / rows32);
    a.n_fast = n_fast < 0.8 * m_fast ? 1 : 0;
    if (banded < 0.8 * (a.n_fast ? n_fast : m_fast) && tokens > rows32 && g_band_mode != 0) band_rows = rows32;
  }
  a.f32_coalesced = (!swap && mode == PG_EPI_F32 && (features % 4) == 0 && (ldo % 4) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                     (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0) &&
                     (resid == nullptr || ((ldr % 4) == 0 && (reinterpret_cast<uintptr_t>(resid) & 15) == 0))) ? 1 : 0;
  a.trace = g_gemm_trace ? g_gemm_trace + 8 * (g_gemm_trace_idx++ % 64) : nullptr;

  CUtensorMap ta, tb;
  int rc;
  if (swap) {
    int BN = tokens <= 16 ? 16 : tokens <= 32 ? 32 : tokens <= 64 ? 64 : 128;
    if ((rc = make_tmap_2d(&ta, w, features, K, ldw, BM)) != PG_OK) return rc;
    if ((rc = make_tmap_2d(&tb, x, tokens, K, ldx, BN)) != PG_OK) return rc;
    const int tiles = ((features + BM - 1) / BM) * split_k;
    // 33..64 tokens, red.add or GEGLU epilogue: the 320-thread variant with eight epilogue warps
    const bool eight = (mode == PG_EPI_ATOMIC_F32 || mode == PG_EPI_GEGLU) && BN == 64;
    if (eight) return launch<64, true, true>(ta, tb, a, tiles, st);
    switch (BN) {
      case 16: return launch<16, true>(ta, tb, a, tiles, st);
      case 32: return launch<32, true>(ta, tb, a, tiles, st);
      case 64: return launch<64, true>(ta, tb, a, tiles, st);
      default: return launch<128, true>(ta, tb, a, tiles, st);
    }
  } else {
    // largest N tile that still gives the persistent grid ~a wave of CTAs (small token counts: latency path)
    const int m_blocks = (tokens + BM - 1) / BM;
    auto ntiles = [&](int bn) { return m_blocks * ((features + bn - 1) / bn) * split_k; };
    int BN;
    if (mode == PG_EPI_GEGLU) {
      BN = (features >= 256 && ntiles(256) >= 120) ? 256 : 128;
    } else {
      BN = features > 128 ? 256 : features > 64 ? 128 : 64;
      while (BN > 64 && ntiles(BN) < 120) BN /= 2;
      if (g_force_bn) BN = g_force_bn;
    }
    if ((rc = make_tmap_2d(&ta, x, tokens, K, ldx, BM)) != PG_OK) return rc;
    // many tokens: a PAIR of CTAs per 256 x 256 tile (cta_group::2), each staging half of the weight tile
    const int pair_tiles = ((tokens + 2 * BM - 1) / (2 * BM)) * ((features + 255) / 256);
    // (measured per call site, profiles/r02h_gemm_pair_ab.txt: +2..16 % everywhere but the gelu epilogue of fc1, -3 %)
    if (g_pair_mode != 0 && BN == 256 && split_k == 1 && pair_tiles >= g_pair_min_tiles && !(mode == PG_EPI_BF16 && act_gelu)) {
      if (band_rows > 0) a.n_fast = 2 + (static_cast<int>(band_rows) / (2 * BM) > 0 ? static_cast<int>(band_rows) / (2 * BM) : 1);
      if ((rc = make_tmap_2d(&tb, w, features, K, ldw, PAIR_BN / 2)) != PG_OK) return rc;
      // eight epilogue warps where the epilogue outlasts the main loop: short K, and the fp32 residual epilogue (8 bytes of
      // residual traffic per element) up to K = 2048 -- o_proj 142 -> 122 us at 16.6k tokens; gate||up and K >= 4304 prefer four
      const bool epi8 = g_epi8_mode == 1 ? true : (g_epi8_mode == 2 ? false : (K <= 1536 || (mode == PG_EPI_F32 && K <= 2048)));
      return epi8 ? launch_pair<true, false>(ta, tb, a, pair_tiles, st) : launch_pair<false, false>(ta, tb, a, pair_tiles, st);
    }
    if (band_rows > 0) a.n_fast = 2 + (static_cast<int>(band_rows) / BM > 0 ? static_cast<int>(band_rows) / BM : 1);
    if ((rc = make_tmap_2d(&tb, w, features, K, ldw, BN)) != PG_OK) return rc;
    const int tiles = ((tokens + BM - 1) / BM) * ((features + BN - 1) / BN) * split_k;
    // short reduction (K <= 1536: the SigLIP projections): the epilogue outlasts the main loop, so it gets eight warps
    if (BN == 256 && K <= 1536) return launch<256, false, true>(ta, tb, a, tiles, st);
    switch (BN) {
      case 64: return launch<64, false>(ta, tb, a, tiles, st);
      case 128: return launch<128, false>(ta, tb, a, tiles, st);
      default: return launch<256, false>(ta, tb, a, tiles, st);
    }
  }
}

extern "C" int pg_gemm_qkv_rope(const void* x, long long ldx, const void* w, long long ldw, void* qkv_out, long long ldo, int tokens,
                                int K, int Hq, int Hkv, int dh, const int* pos, const float* inv_freq, void* k_pages, void* v_pages,
                                const int* page_table, const int* slot_base, int tokens_per_seq, int page_size, int max_pages,
                                void* stream) {
  if (tokens <= 0 || K <= 0 || Hq <= 0 || Hkv <= 0 || tokens_per_seq <= 0 || pos == nullptr || inv_freq == nullptr) return PG_ERR_ARG;
  if (dh != 64 && dh != 256) return PG_ERR_ARG;  // one head per N tile: instantiated for head_dim 64 and 256
  if ((K % 8) != 0 || (ldx % 8) != 0 || (ldw % 8) != 0 || (ldo % 8) != 0) return PG_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(qkv_out) & 15))
    return PG_ERR_ARG;
  const bool cache = k_pages != nullptr;
  if (cache && (v_pages == nullptr || page_table == nullptr || slot_base == nullptr || page_size != 64 || max_pages <= 0 ||
                (reinterpret_cast<uintptr_t>(k_pages) & 15) || (reinterpret_cast<uintptr_t>(v_pages) & 15)))
    return PG_ERR_ARG;
  const int features = (Hq + 2 * Hkv) * dh;
  GemmArgs a = {};
  a.tokens = tokens; a.features = features; a.K = K; a.split_k = 1; a.mode = PG_EPI_BF16; a.scale = 1.f;
  a.out = qkv_out; a.ldo = ldo;
  a.rope_pos = pos; a.rope_inv_freq = inv_freq; a.rope_hq = Hq; a.rope_hkv = Hkv;
  a.k_pages = static_cast<__nv_bfloat16*>(k_pages); a.v_pages = static_cast<__nv_bfloat16*>(v_pages);
  a.page_table = page_table; a.slot_base = slot_base; a.tokens_per_seq = tokens_per_seq; a.max_pages = max_pages;
  a.n_fast = 1;  // every head of a few token tiles at a time: the activations are read once (token tiles first re-read them once per head)
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap_2d(&ta, x, tokens, K, ldx, BM)) != PG_OK) return rc;
  if ((rc = make_tmap_2d(&tb, w, features, K, ldw, dh)) != PG_OK) return rc;
  const int tiles = ((tokens + BM - 1) / BM) * (Hq + 2 * Hkv);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int pair_tiles = ((tokens + 2 * BM - 1) / (2 * BM)) * (Hq + 2 * Hkv);
  if (dh == 256 && g_pair_mode != 0 && pair_tiles >= g_pair_min_tiles) {  // CTA pair: 256 tokens x one head per tile
    if ((rc = make_tmap_2d(&tb, w, features, K, ldw, PAIR_BN / 2)) != PG_OK) return rc;
    return launch_pair<false, true>(ta, tb, a, pair_tiles, st);
  }
  return dh == 256 ? launch<256, false, false, true>(ta, tb, a, tiles, st) : launch<64, false, false, true>(ta, tb, a, tiles, st);
}
