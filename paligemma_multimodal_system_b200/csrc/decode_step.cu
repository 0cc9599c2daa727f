// One persistent kernel for a whole Gemma decode step (q_len = 1, <= 64 sequences): one CTA per SM, every layer's
//
//   RMSNorm -> QKV GEMM -> RoPE + KV append + attention -> O GEMM -> RMSNorm -> gate||up GEGLU GEMM -> down GEMM
//
// runs as a PHASE of the same grid, separated by grid-wide barriers instead of kernel boundaries, followed by the final
// RMSNorm and the lm_head GEMM (modeling_gemma.py:385-418,453-472,501-533 at q_len == 1).  Per CTA:
//   warp 0      TMA producer: streams the 128-row weight slabs of this CTA's GEMM tiles through a 6-stage ring; the
//               weights of the NEXT GEMM phase are fetched while the grid is still in a barrier / norm phase
//   warp 1      tcgen05.mma issuer, fp32 accumulators double buffered in TMEM (allocated once for the whole step)
//   warps 2-9   workers: GEMM epilogues (warps 2-5, one TMEM lane quadrant each), RMSNorm rows, attention
// Split-K partial sums are red.add-ed into the fp32 residual stream / the fp32 qkv buffer exactly as in the
// multi-kernel path (gemm_tcgen05.cu), so both paths produce the same numbers up to fp32 summation order.
#include <cooperative_groups.h>

#include "common.cuh"
#include "paligemma_b200.h"

namespace pg {
namespace ds {

typedef __nv_bfloat16 bf16;

constexpr int BM = 128, BK = 64;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB weight tile
constexpr int STAGES = 6;
constexpr int NWORK = 256;            // worker threads (warps 2..9)
constexpr int NTHREADS = 64 + NWORK;
constexpr int PAGE = 64;
constexpr int NBUF = 3;               // attention page ring (aliases the GEMM stage memory)

enum { M_QKV = 0, M_O = 1, M_GU = 2, M_DOWN = 3 };  // tensor-map slots per layer; then HEAD, then B-operand maps

struct Params {
  const CUtensorMap* maps;  // [4*L] layer weights, [4L] head, [4L+1] hn, [4L+2] att, [4L+3] mid   (global memory)
  int L, B, D, F, Hq, Hkv, dh, V, W;
  int split_qkv, split_o, split_down;
  // decode-step inputs
  const int* cur_tok;
  const bf16* embed;
  const float* img;
  int n_img;
  float text_scale, img_scale;
  long long pad_token, image_token;
  // activations
  float* h;
  bf16* hn;
  float* qkv;
  bf16* att;
  bf16* mid;
  float* logits;
  // norm weights (fp32): ln1 [L, D], ln2 [L, D], final [D]; lm_head bias [V]
  const float* ln1;
  const float* ln2;
  const float* norm_w;
  const float* head_b;
  float eps;
  // paged KV cache
  bf16* k_pages;
  bf16* v_pages;
  long long layer_stride;  // elements between consecutive layers' page pools
  const int* page_table;
  const int* pos;
  const int* kv_len;
  const float* inv_freq;
  int max_pages;
  float sl2;
  // grid barrier state (persists across launches)
  unsigned int* bar_flags;  // [gridDim.x] epoch published by each CTA (persists across launches)
  unsigned long long* trace;  // optional: globaltimer (ns) of CTA trace_cta after every phase barrier + SM-clock detail
  int trace_cta;
};

template <int BN>
struct Smem {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int XCH_BYTES = 64 * BN * 4;
};

PG_DEVINL void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
PG_DEVINL unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
PG_DEVINL int ld_acquire_cta_smem(const volatile int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(const_cast<const int*>(p))) : "memory");
  return v;
}
PG_DEVINL void st_release_cta_smem(volatile int* p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(const_cast<int*>(p))), "r"(v) : "memory");
}

PG_DEVINL unsigned ld_relaxed_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
PG_DEVINL void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
PG_DEVINL void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Grid barrier without atomics: CTA c publishes the epoch in flags[c] (release), one warp per CTA polls all flags
// (coalesced: 148 flags = 5 cache lines).  Called by ONE WARP per CTA after a CTA-level barrier has ordered the
// CTA's writes before it; every CTA of the grid must call it the same number of times.
PG_DEVINL void grid_barrier(unsigned* flags, unsigned nctas, unsigned cta, unsigned epoch, int lane) {
  if (lane == 0) st_release_gpu(flags + cta, epoch);  // release is cumulative over the writes ordered before it
  long long t0 = clock64();
  while (true) {
    bool ok = true;
    for (unsigned i = lane; i < nctas; i += 32) ok &= (ld_relaxed_gpu(flags + i) - epoch) < 0x80000000u;  // flag >= epoch
    if (__all_sync(0xffffffffu, ok)) break;
    if (clock64() - t0 > 4000000000LL) {
      if (lane == 0) printf("pg: grid barrier timeout block %u epoch %u\n", cta, epoch);
      __trap();
    }
  }
  fence_acq_rel_gpu();
}

struct GemmPhase {
  int map_a, map_b;   // tensor-map indices
  int features, K, split_k, mode;
  void* out;
  long long ldo;
  const float* bias;
};

PG_DEVINL void tile_of(int item, int m_blocks, int total_kb, int split_k, int& m_blk, int& kb0, int& kb1) {
  m_blk = item % m_blocks;
  const int split = item / m_blocks;
  const int per = (total_kb + split_k - 1) / split_k;
  kb0 = split * per;
  kb1 = min(total_kb, kb0 + per);
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1) decode_step_kernel(const __grid_constant__ Params p) {
  using S = Smem<BN>;
  constexpr uint32_t IDESC = make_idesc_bf16(BM, BN);
  constexpr int TMEM_COLS = (2 * BN) <= 32 ? 32 : (2 * BN) <= 64 ? 64 : 128;

  extern __shared__ __align__(1024) uint8_t smem[];
  // [0, STAGES*STAGE_BYTES)  GEMM ring   |  then XCH   -- the attention page ring aliases both (and more)
  const uint32_t smem_base = smem_u32(smem);
  if ((smem_base & 1023u) != 0) __trap();
  float* xch = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES);
  // attention view of the same memory
  const int LDS = p.dh + 8;
  const int buf_elems = 2 * PAGE * LDS;
  bf16* ring = reinterpret_cast<bf16*>(smem);
  // fixed area after the aliased region
  uint8_t* fixed = smem + max(STAGES * S::STAGE_BYTES + S::XCH_BYTES, NBUF * buf_elems * 2);
  fixed = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fixed) + 127) & ~uintptr_t(127));
  bf16* Qs = reinterpret_cast<bf16*>(fixed);                              // [16][LDS]
  bf16* new_k = Qs + 16 * LDS;                                            // [dh]
  bf16* new_v = new_k + p.dh;                                             // [dh]
  float* scratch = reinterpret_cast<float*>(new_v + p.dh);                // 64 floats (block reductions)
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + 64);             // full[S], empty[S], tfull[2], tempty[2], kv[NBUF]
  volatile int* phase_done = reinterpret_cast<volatile int*>(bars + 2 * STAGES + 4 + NBUF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_cast<int*>(phase_done) + 1);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
  auto kv_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 4 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  const int total_phases_marker = 0;
  (void)total_phases_marker;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
    for (int s = 0; s < NBUF; ++s) mbar_init(kv_bar(s), 1);
    *phase_done = 0;
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();

  // ---- the static schedule of GEMM phases (identical in every role) ----
  const int n_gemm = 4 * p.L + 1;
  auto gemm_phase = [&](int g) -> GemmPhase {
    GemmPhase ph;
    const int mb = 4 * p.L + 1;  // first B-operand map
    if (g == 4 * p.L) {
      ph = {4 * p.L, mb + 0, p.V, p.D, 1, PG_EPI_F32, p.logits, p.V, p.head_b};
    } else {
      const int l = g >> 2, k = g & 3;
      if (k == M_QKV) ph = {g, mb + 0, p.W, p.D, p.split_qkv, PG_EPI_ATOMIC_F32, p.qkv, p.W, nullptr};
      else if (k == M_O) ph = {g, mb + 1, p.D, p.D, p.split_o, PG_EPI_ATOMIC_F32, p.h, p.D, nullptr};
      else if (k == M_GU) ph = {g, mb + 0, 2 * p.F, p.D, 1, PG_EPI_GEGLU, p.mid, p.F, nullptr};
      else ph = {g, mb + 2, p.D, p.F, p.split_down, PG_EPI_ATOMIC_F32, p.h, p.D, nullptr};
      (void)l;
    }
    return ph;
  };
  // number of grid barriers that must have completed before GEMM phase g may read its activations / may touch the
  // (attention-aliased) stage memory.  Barrier order per layer: [A norm1][B qkv][C attn][D o][E norm2][F gu][G down]
  auto barriers_before_b = [&](int g) -> int {
    if (g == 4 * p.L) return 7 * p.L + 1;  // after the final norm
    const int l = g >> 2, k = g & 3;
    return 7 * l + (k == M_QKV ? 1 : k == M_O ? 3 : k == M_GU ? 5 : 6);
  };
  auto barriers_before_a = [&](int g) -> int {  // the O GEMM must not prefetch while attention owns the memory
    if (g == 4 * p.L) return 0;
    return ((g & 3) == M_O) ? 7 * (g >> 2) + 3 : 0;
  };

  if (warp == 0) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto wait_phase = [&](int need) {
        if (need <= 0) return;
        long long t0 = clock64();
        while (ld_acquire_cta_smem(phase_done) < need) {
          if (clock64() - t0 > 4000000000LL) { printf("pg: producer phase wait timeout block %d need %d\n", cta, need); __trap(); }
        }
      };
      for (int g = 0; g < n_gemm; ++g) {
        const GemmPhase ph = gemm_phase(g);
        const CUtensorMap* ma = p.maps + ph.map_a;
        const CUtensorMap* mb = p.maps + ph.map_b;
        const int m_blocks = (ph.features + BM - 1) / BM;
        const int total_kb = (ph.K + BK - 1) / BK;
        const int items = m_blocks * ph.split_k;
        wait_phase(barriers_before_a(g));
        // weights first (they never depend on other CTAs): up to STAGES tiles in flight before the activations exist
        int pre = 0;
        {
          int st = stage;
          uint32_t phs = phase;
          for (int item = cta; item < items && pre < STAGES; item += G) {
            int m_blk, kb0, kb1;
            tile_of(item, m_blocks, total_kb, ph.split_k, m_blk, kb0, kb1);
            for (int kb = kb0; kb < kb1 && pre < STAGES; ++kb) {
              mbar_wait(empty_bar(st), phs ^ 1);
              mbar_expect_tx(full_bar(st), S::STAGE_BYTES);
              tma_load_2d(smem_base + st * S::STAGE_BYTES, ma, full_bar(st), kb * BK, m_blk * BM, kEvictFirst);
              ++pre;
              if (++st == STAGES) { st = 0; phs ^= 1; }
            }
          }
        }
        wait_phase(barriers_before_b(g));
        for (int item = cta; item < items; item += G) {
          int m_blk, kb0, kb1;
          tile_of(item, m_blocks, total_kb, ph.split_k, m_blk, kb0, kb1);
          for (int kb = kb0; kb < kb1; ++kb) {
            const uint32_t sa = smem_base + stage * S::STAGE_BYTES;
            if (pre > 0) {
              --pre;
            } else {
              mbar_wait(empty_bar(stage), phase ^ 1);
              mbar_expect_tx(full_bar(stage), S::STAGE_BYTES);
              tma_load_2d(sa, ma, full_bar(stage), kb * BK, m_blk * BM, kEvictFirst);
            }
            tma_load_2d(sa + A_BYTES, mb, full_bar(stage), kb * BK, 0, kEvictLast);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();  // the whole warp reaches the CTA-wide barrier below together (bar.sync counts warps, not lanes)
  } else if (warp == 1) {
    // =========================================== MMA issuer =============================================
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int g = 0; g < n_gemm; ++g) {
        const GemmPhase ph = gemm_phase(g);
        const int m_blocks = (ph.features + BM - 1) / BM;
        const int total_kb = (ph.K + BK - 1) / BK;
        const int items = m_blocks * ph.split_k;
        for (int item = cta; item < items; item += G) {
          int m_blk, kb0, kb1;
          tile_of(item, m_blocks, total_kb, ph.split_k, m_blk, kb0, kb1);
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * S::STAGE_BYTES;
            const uint64_t adesc = make_sdesc_k_sw128(sa);
            const uint64_t bdesc = make_sdesc_k_sw128(sa + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit(empty_bar(stage));
            if (kb == kb1 - 1) umma_commit(tfull_bar(acc));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================================== workers ================================================
    const int wtid = threadIdx.x - 64;       // 0..255
    const int wwarp = wtid >> 5;             // 0..7
    const int q = warp & 3;                  // TMEM lane quadrant (epilogue warps are worker warps 0..3)
    int acc = 0;
    uint32_t acc_phase = 0;
    int n_barriers = 0;
    uint32_t kv_uses[NBUF] = {0, 0, 0};
    // epoch base: the value this CTA published last (equal in every CTA at a kernel boundary)
    unsigned epoch = ld_acquire_gpu(p.bar_flags + cta);
    if (wtid == 0 && p.trace != nullptr && cta == p.trace_cta) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.trace[0] = t;
    }

    auto phase_barrier = [&]() {
      const long long c_work = clock64();
      named_bar(2, NWORK);  // every worker's global writes (stores, red.add) are ordered before the release below
      const long long c_local = clock64();
      ++epoch;
      ++n_barriers;
      if (wwarp == 0) {
        grid_barrier(p.bar_flags, G, cta, epoch, lane);
        if (lane == 0) {
          st_release_cta_smem(phase_done, n_barriers);
          if (p.trace != nullptr && cta == p.trace_cta) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.trace[n_barriers] = t;
            // SM-clock detail: [work done by thread 0][all workers of this CTA done][grid barrier passed]
            p.trace[1024 + 3 * n_barriers + 0] = c_work;
            p.trace[1024 + 3 * n_barriers + 1] = c_local;
            p.trace[1024 + 3 * n_barriers + 2] = clock64();
          }
        }
      }
      named_bar(2, NWORK);
    };
    auto block_sum = [&](float v) -> float {  // over the 256 workers
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      named_bar(2, NWORK);
      if (lane == 0) scratch[wwarp] = v;
      named_bar(2, NWORK);
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += scratch[w];
      return t;
    };
    // RMSNorm of residual row `cta` (GemmaRMSNorm, modeling_gemma.py:172-181); optionally gathers the token embedding
    // first (layer 0) and zeroes the fp32 qkv row for the split-K QKV GEMM
    auto norm_phase = [&](const float* w, bool embed_first, bool zero_qkv) {
      if (cta < p.B) {
        float* hr = p.h + static_cast<long long>(cta) * p.D;
        float ss = 0.f;
        if (embed_first) {
          const long long id = p.cur_tok[cta];
          for (int i = wtid; i < p.D; i += NWORK) {
            float v;
            if (id == p.pad_token) v = 0.f;
            else if (id == p.image_token && p.img != nullptr) v = p.img[static_cast<long long>(cta) * p.n_img * p.D + i] * p.img_scale;
            else v = __bfloat162float(p.embed[id * p.D + i]) * p.text_scale;
            hr[i] = v;
            ss += v * v;
          }
        } else {
          for (int i = wtid; i < p.D; i += NWORK) {
            const float v = __ldcg(hr + i);
            ss += v * v;
          }
        }
        const float rstd = rsqrtf(block_sum(ss) / p.D + p.eps);
        bf16* o = p.hn + static_cast<long long>(cta) * p.D;
        for (int i = wtid; i < p.D; i += NWORK) o[i] = __float2bfloat16(__ldcg(hr + i) * rstd * (1.0f + w[i]));
        if (zero_qkv) {
          float* z = p.qkv + static_cast<long long>(cta) * p.W;
          for (int i = wtid; i < p.W; i += NWORK) z[i] = 0.f;
        }
      }
    };
    // GEMM epilogue of this CTA's tiles (weight rows on the TMEM lanes, tokens on the columns)
    auto gemm_epilogue = [&](int g) {
      const GemmPhase ph = gemm_phase(g);
      const int m_blocks = (ph.features + BM - 1) / BM;
      const int total_kb = (ph.K + BK - 1) / BK;
      const int items = m_blocks * ph.split_k;
      if (wwarp >= 4) return;  // only the four epilogue warps touch TMEM
      bf16* out_bf = reinterpret_cast<bf16*>(ph.out);
      float* out_f = reinterpret_cast<float*>(ph.out);
      for (int item = cta; item < items; item += G) {
        int m_blk, kb0, kb1;
        tile_of(item, m_blocks, total_kb, ph.split_k, m_blk, kb0, kb1);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
        const int rl = q * 32 + lane;
        if (ph.mode == PG_EPI_GEGLU) {
          if (q >= 2) {
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
              if (c0 >= p.B) break;
              uint32_t r[16];
              tmem_ld16(taddr + c0, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) xch[(rl - 64) * BN + ((c0 + i) ^ ((rl - 64) & (BN - 1) & 31))] = __uint_as_float(r[i]);
            }
          }
          named_bar(1, 128);
          if (q < 2) {
            const int fo = m_blk * 64 + rl;
            const bool f_ok = fo < ph.features / 2;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
              if (c0 >= p.B) break;
              uint32_t r[16];
              tmem_ld16(taddr + c0, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int j = c0 + i;
                if (f_ok && j < p.B) {
                  const float val = gelu_tanh(__uint_as_float(r[i])) * xch[rl * BN + ((c0 + i) ^ (rl & (BN - 1) & 31))];
                  out_bf[static_cast<long long>(j) * ph.ldo + fo] = __float2bfloat16(val);
                }
              }
            }
          }
          named_bar(1, 128);
        } else {
          const int fr = m_blk * BM + rl;
          const bool f_ok = fr < ph.features;
          const float bias = (ph.bias != nullptr && f_ok && kb0 == 0) ? __ldg(ph.bias + fr) : 0.f;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 16) {
            if (c0 >= p.B) break;
            uint32_t r[16];
            tmem_ld16(taddr + c0, r);
            tmem_ld_wait();
            if (!f_ok) continue;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int j = c0 + i;
              if (j < p.B) {
                const float x = __uint_as_float(r[i]) + bias;
                const long long o = static_cast<long long>(j) * ph.ldo + fr;
                if (ph.mode == PG_EPI_F32) out_f[o] = x;
                else atomicAdd(out_f + o, x);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    };
    // RoPE + KV append + attention of sequence `cta` for layer l (same math as attn_decode_fused_kernel, one CTA per
    // (sequence, kv head), 8 worker warps = two groups of 4 working on alternate pages)
    auto attention_phase = [&](int l) {
      if (cta >= p.B * p.Hkv) return;
      const int DH = p.dh, HALF = DH / 2, group = p.Hq / p.Hkv;
      const int b = cta / p.Hkv, hk = cta % p.Hkv;
      const int wg = wwarp >> 2, wq = wwarp & 3;  // (output-dh half, 16-key quarter of a page)
      bf16* kp = p.k_pages + static_cast<long long>(l) * p.layer_stride;
      bf16* vp = p.v_pages + static_cast<long long>(l) * p.layer_stride;
      const int len = p.kv_len[b];
      const int n_tiles = (len + PAGE - 1) / PAGE;
      const int new_slot = len - 1, new_tile = new_slot / PAGE;
      const long long kv_ts = static_cast<long long>(p.Hkv) * DH;
      const int* ptab = p.page_table + b * p.max_pages;
      const uint32_t row_bytes = DH * 2;
      auto load_kv = [&](int tile, int slot) {
        const int page = ptab[tile];
        const int n0 = tile * PAGE;
        int rows = min(PAGE, len - n0);
        if (tile == new_tile) rows -= 1;
        if (wtid == 0) mbar_expect_tx(kv_bar(slot), static_cast<uint32_t>(rows) * row_bytes * 2);
        if (wtid < 2 * PAGE) {
          const int kv = wtid >> 6, r = wtid & 63;
          bf16* dst = ring + slot * buf_elems + kv * PAGE * LDS + r * LDS;
          if ((n0 + r) < len && (n0 + r) != new_slot) {
            const bf16* src = (kv ? vp : kp) + (static_cast<long long>(page) * PAGE + r) * kv_ts + hk * DH;
            bulk_copy_g2s(smem_u32(dst), src, row_bytes, kv_bar(slot));
          } else {
            for (int c = 0; c < DH / 8; ++c) reinterpret_cast<uint4*>(dst)[c] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      };
      for (int i = 0; i < NBUF && i < n_tiles; ++i) load_kv(i, i);
      // RoPE of q (and of the new k), staged in shared memory
      const float* __restrict__ row = p.qkv + static_cast<long long>(b) * p.W;
      const float posf = static_cast<float>(p.pos[b]);
      for (int i = wtid; i < HALF; i += NWORK) {
        float x1[16], x2[16];
#pragma unroll
        for (int g2 = 0; g2 < 16; ++g2) {
          if (g2 < group) {
            const float* qh = row + (hk * group + g2) * DH;
            x1[g2] = __ldcg(qh + i);
            x2[g2] = __ldcg(qh + i + HALF);
          }
        }
        const float* kh = row + (p.Hq + hk) * DH;
        const float* vh = row + (p.Hq + p.Hkv + hk) * DH;
        const float kx1 = __ldcg(kh + i), kx2 = __ldcg(kh + i + HALF), vx1 = __ldcg(vh + i), vx2 = __ldcg(vh + i + HALF);
        float sn, cs;
        sincosf(posf * p.inv_freq[i], &sn, &cs);
#pragma unroll
        for (int g2 = 0; g2 < 16; ++g2) {
          if (g2 < group) {
            Qs[g2 * LDS + i] = __float2bfloat16(x1[g2] * cs - x2[g2] * sn);
            Qs[g2 * LDS + i + HALF] = __float2bfloat16(x2[g2] * cs + x1[g2] * sn);
          }
        }
        const bf16 k1 = __float2bfloat16(kx1 * cs - kx2 * sn), k2 = __float2bfloat16(kx2 * cs + kx1 * sn);
        const bf16 v1 = __float2bfloat16(vx1), v2 = __float2bfloat16(vx2);
        const int page = ptab[new_tile];
        bf16* kb = kp + (static_cast<long long>(page) * PAGE + (new_slot - new_tile * PAGE)) * kv_ts + hk * DH;
        bf16* vb = vp + (static_cast<long long>(page) * PAGE + (new_slot - new_tile * PAGE)) * kv_ts + hk * DH;
        kb[i] = k1; kb[i + HALF] = k2;
        vb[i] = v1; vb[i + HALF] = v2;
        new_k[i] = k1; new_k[i + HALF] = k2;
        new_v[i] = v1; new_v[i + HALF] = v2;
      }
      for (int idx = wtid; idx < 16 * DH; idx += NWORK) {
        const int r = idx / DH;
        if (r >= group) Qs[r * LDS + idx % DH] = __float2bfloat16(0.f);
      }
      named_bar(2, NWORK);
      if (new_tile < NBUF) {
        bf16* Kb = ring + new_tile * buf_elems;
        const int r = new_slot - new_tile * PAGE;
        for (int k = wtid; k < DH; k += NWORK) {
          Kb[r * LDS + k] = new_k[k];
          Kb[(PAGE + r) * LDS + k] = new_v[k];
        }
      }
      named_bar(2, NWORK);

      // Work split: warp quarter wq owns 16 keys of the page; warp group wg owns one half of the head dimension of the
      // output (both groups compute the cheap QK^T scores; splitting O keeps the accumulators at 64 registers).
      float o[16][4];
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
      float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
      const uint32_t q_addr = smem_u32(Qs + (lane & 15) * LDS + (lane >> 4) * 8);
      const int ksteps = DH / 16;      // 16 (dh 256) or 4 (dh 64)
      const int dps = ksteps / 2;      // 16-column blocks of O per warp group
      for (int i = 0; i < n_tiles; ++i) {
        const int slot = i % NBUF;
        mbar_wait(kv_bar(slot), kv_uses[slot] & 1);
        const bf16* Kt = ring + slot * buf_elems + wq * 16 * LDS;
        const bf16* Vt = Kt + PAGE * LDS;
        float s[2][4], s2[2][4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[a][e] = s2[a][e] = 0.f;
        const uint32_t k_addr = smem_u32(Kt + ((lane & 7) + (lane >> 4) * 8) * LDS + ((lane >> 3) & 1) * 8);
#pragma unroll 4
        for (int ks = 0; ks < ksteps; ks += 2) {
          uint32_t a[4], b0, b1, b2, b3, c4[4], d0, d1, d2, d3;
          ldmatrix_x4(q_addr + ks * 32, a[0], a[1], a[2], a[3]);
          ldmatrix_x4(k_addr + ks * 32, b0, b1, b2, b3);
          ldmatrix_x4(q_addr + (ks + 1) * 32, c4[0], c4[1], c4[2], c4[3]);
          ldmatrix_x4(k_addr + (ks + 1) * 32, d0, d1, d2, d3);
          mma_bf16_16816(s[0], a, b0, b1);
          mma_bf16_16816(s[1], a, b2, b3);
          mma_bf16_16816(s2[0], c4, d0, d1);
          mma_bf16_16816(s2[1], c4, d2, d3);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[a][e] += s2[a][e];
        const int kbase = i * PAGE + wq * 16;
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
        if (i > 0) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            o[k][0] *= alpha[0]; o[k][1] *= alpha[0];
            o[k][2] *= alpha[1]; o[k][3] *= alpha[1];
          }
        }
        uint32_t a[4];
        a[0] = pack_bf16(s[0][0], s[0][1]);
        a[1] = pack_bf16(s[0][2], s[0][3]);
        a[2] = pack_bf16(s[1][0], s[1][1]);
        a[3] = pack_bf16(s[1][2], s[1][3]);
        const uint32_t v_addr = smem_u32(Vt + ((lane & 7) + ((lane >> 3) & 1) * 8) * LDS + (wg * dps * 16) + (lane >> 4) * 8);
#pragma unroll
        for (int dp = 0; dp < 8; ++dp) {
          if (dp < dps) {
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(v_addr + dp * 32, b0, b1, b2, b3);
            mma_bf16_16816(o[2 * dp], a, b0, b1);
            mma_bf16_16816(o[2 * dp + 1], a, b2, b3);
          }
        }
        named_bar(2, NWORK);  // every warp is done with this slot
        kv_uses[slot] += 1;
        if (i + NBUF < n_tiles) {  // refill (long contexts only)
          fence_proxy_async_smem();
          load_kv(i + NBUF, slot);
          if (i + NBUF == new_tile) {
            named_bar(2, NWORK);
            bf16* Kb = ring + slot * buf_elems;
            const int r = new_slot - new_tile * PAGE;
            for (int qq = wtid; qq < DH; qq += NWORK) {
              Kb[r * LDS + qq] = new_k[qq];
              Kb[(PAGE + r) * LDS + qq] = new_v[qq];
            }
          }
          named_bar(2, NWORK);
        }
      }
      named_bar(2, NWORK);
      // merge the 4 key quarters through shared memory (rows < group only; group <= 8 on this path)
      const int RLD = DH + 2;
      float* red = reinterpret_cast<float*>(ring);  // [4 quarters][8 rows][RLD]
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
      }
      {
        float* dst0 = red + (wq * 8 + (lane >> 2)) * RLD + wg * dps * 16;
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
          if (nt < 2 * dps) {
            const int col = nt * 8 + (lane & 3) * 2;
            dst0[col] = o[nt][0];
            dst0[col + 1] = o[nt][1];
          }
        }
        if (wg == 0 && (lane & 3) == 0) {
          float* ml = red + (wq * 8 + (lane >> 2)) * RLD + DH;
          ml[0] = m_run[0] * p.sl2;
          ml[1] = l_run[0];
        }
      }
      named_bar(2, NWORK);
      const long long hq0 = static_cast<long long>(b) * p.Hq + hk * group;
      for (int idx = wtid; idx < group * DH; idx += NWORK) {
        const int r = idx / DH, col = idx % DH;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 4; ++w) M = fmaxf(M, red[(w * 8 + r) * RLD + DH]);
        float accv = 0.f, Lsum = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float wgt = exp2f(red[(w * 8 + r) * RLD + DH] - M);
          accv += red[(w * 8 + r) * RLD + col] * wgt;
          Lsum += red[(w * 8 + r) * RLD + DH + 1] * wgt;
        }
        p.att[(hq0 + r) * DH + col] = __float2bfloat16(accv / Lsum);
      }
      named_bar(2, NWORK);
      fence_proxy_async_smem();  // the stage memory goes back to the TMA producer after the barrier
    };

    for (int l = 0; l < p.L; ++l) {
      norm_phase(p.ln1 + static_cast<long long>(l) * p.D, l == 0, true);
      phase_barrier();
      gemm_epilogue(4 * l + M_QKV);
      phase_barrier();
      attention_phase(l);
      phase_barrier();
      gemm_epilogue(4 * l + M_O);
      phase_barrier();
      norm_phase(p.ln2 + static_cast<long long>(l) * p.D, false, false);
      phase_barrier();
      gemm_epilogue(4 * l + M_GU);
      phase_barrier();
      gemm_epilogue(4 * l + M_DOWN);
      phase_barrier();
    }
    norm_phase(p.norm_w, false, false);
    phase_barrier();
    gemm_epilogue(4 * p.L);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) griddep_launch_dependents();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN>
static size_t smem_bytes_for(int dh) {
  using S = Smem<BN>;
  const size_t aliased = max(static_cast<size_t>(STAGES * S::STAGE_BYTES + S::XCH_BYTES), static_cast<size_t>(NBUF) * 2 * PAGE * (dh + 8) * 2);
  const size_t fixed = 128 + static_cast<size_t>(16) * (dh + 8) * 2 + 2 * dh * 2 + 64 * 4 + (2 * STAGES + 4 + NBUF) * 8 + 16;
  return aliased + fixed + 128;
}

template <int BN>
static int launch(const Params& p, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = smem_bytes_for<BN>(p.dh);
  if (!configured) {
    if (cudaFuncSetAttribute(decode_step_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
      cudaGetLastError();
      return PG_ERR_CUDA;
    }
    configured = true;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;  // every CTA must be resident: the phases are separated by grid barriers
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pg_pdl_enabled() ? 2 : 1;
  pg_count_launch(1);
  return cudaLaunchKernelEx(&cfg, decode_step_kernel<BN>, p) == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}

}  // namespace ds
}  // namespace pg

using namespace pg;

// see include/paligemma_b200.h
extern "C" int pg_decode_step(const PgDecodeStepArgs* a, void* stream) {
  if (a == nullptr) return PG_ERR_ARG;
  if (a->B <= 0 || a->B > 64 || a->L <= 0 || (a->dh != 256 && a->dh != 64) || a->Hkv <= 0 || a->Hq % a->Hkv != 0 ||
      a->Hq / a->Hkv > 8 || a->B * a->Hkv > 148 || (a->F % 64) != 0 || (a->D % 8) != 0 || a->page_size != 64)
    return PG_ERR_ARG;
  ds::Params p;
  p.maps = static_cast<const CUtensorMap*>(a->tensor_maps);
  p.L = a->L; p.B = a->B; p.D = a->D; p.F = a->F; p.Hq = a->Hq; p.Hkv = a->Hkv; p.dh = a->dh; p.V = a->V;
  p.W = (a->Hq + 2 * a->Hkv) * a->dh;
  auto norm_split = [](int split, int K) {  // no empty splits (an empty item would never signal its accumulator)
    const int total_kb = (K + 63) / 64;
    if (split < 1) split = 1;
    if (split > total_kb) split = total_kb;
    const int per = (total_kb + split - 1) / split;
    return (total_kb + per - 1) / per;
  };
  p.split_qkv = norm_split(a->split_qkv, a->D);
  p.split_o = norm_split(a->split_o, a->D);
  p.split_down = norm_split(a->split_down, a->F);
  p.cur_tok = a->cur_tok; p.embed = static_cast<const __nv_bfloat16*>(a->embed); p.img = a->img; p.n_img = a->n_img;
  p.text_scale = a->text_scale; p.img_scale = a->img_scale; p.pad_token = a->pad_token; p.image_token = a->image_token;
  p.h = a->h; p.hn = static_cast<__nv_bfloat16*>(a->hn); p.qkv = a->qkv; p.att = static_cast<__nv_bfloat16*>(a->att);
  p.mid = static_cast<__nv_bfloat16*>(a->mid); p.logits = a->logits;
  p.ln1 = a->ln1; p.ln2 = a->ln2; p.norm_w = a->norm_w; p.head_b = a->head_b; p.eps = a->eps;
  p.k_pages = static_cast<__nv_bfloat16*>(a->k_pages); p.v_pages = static_cast<__nv_bfloat16*>(a->v_pages);
  p.layer_stride = a->layer_stride; p.page_table = a->page_table; p.pos = a->pos; p.kv_len = a->kv_len;
  p.inv_freq = a->inv_freq; p.max_pages = a->max_pages; p.sl2 = a->scale * 1.4426950408889634f;
  p.bar_flags = a->barrier_state;
  p.trace = a->trace;
  p.trace_cta = a->trace_cta;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->B <= 16) return ds::launch<16>(p, st);
  if (a->B <= 32) return ds::launch<32>(p, st);
  return ds::launch<64>(p, st);
}

// Encodes the tensor maps pg_decode_step needs into `out_maps_host` (host memory, (4*L + 4) * 128 bytes); the caller
// copies them to device memory once.  weights[4*l + {0,1,2,3}] = qkv_w, o_w, gu_w (packed), down_w of layer l;
// weights[4*L] = lm_head weight.
extern "C" int pg_decode_step_encode_maps(void* out_maps_host, const void* const* weights, const void* hn, const void* att,
                                          const void* mid, int L, int B, int D, int F, int Hq, int Hkv, int dh, int V) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess || fp == nullptr) return PG_ERR_DRIVER;
  EncodeFn enc = reinterpret_cast<EncodeFn>(fp);
  CUtensorMap* maps = static_cast<CUtensorMap*>(out_maps_host);
  const int W = (Hq + 2 * Hkv) * dh;
  const int BN = B <= 16 ? 16 : B <= 32 ? 32 : 64;
  auto make = [&](CUtensorMap* m, const void* ptr, long long rows, long long K, int box_rows) -> int {
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(K) * 2};
    cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
               ? PG_OK : PG_ERR_TMAP;
  };
  int rc;
  for (int l = 0; l < L; ++l) {
    if ((rc = make(maps + 4 * l + 0, weights[4 * l + 0], W, D, 128)) != PG_OK) return rc;
    if ((rc = make(maps + 4 * l + 1, weights[4 * l + 1], D, D, 128)) != PG_OK) return rc;
    if ((rc = make(maps + 4 * l + 2, weights[4 * l + 2], 2ll * F, D, 128)) != PG_OK) return rc;
    if ((rc = make(maps + 4 * l + 3, weights[4 * l + 3], D, F, 128)) != PG_OK) return rc;
  }
  if ((rc = make(maps + 4 * L, weights[4 * L], V, D, 128)) != PG_OK) return rc;
  if ((rc = make(maps + 4 * L + 1, hn, B, D, BN)) != PG_OK) return rc;
  if ((rc = make(maps + 4 * L + 2, att, B, static_cast<long long>(Hq) * dh, BN)) != PG_OK) return rc;
  if ((rc = make(maps + 4 * L + 3, mid, B, F, BN)) != PG_OK) return rc;
  return PG_OK;
}
