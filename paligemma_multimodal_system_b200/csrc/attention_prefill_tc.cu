// Prefill attention on the 5th-generation tensor cores: non-causal softmax(Q K^T * scale) V, flash style, with both
// GEMMs issued as tcgen05.mma (accumulators in TMEM) and all operands staged by TMA.
//
// Replaces the attention core of the reference at prefill: SigLIP MHA (modeling_siglip.py:96-136; dh = 72, fp32 softmax)
// and the Gemma prefill MQA/GQA (modeling_gemma.py:307-339 with the all-zero mask of modeling_paligemma.py:154-156;
// dh = 256; the G query heads of one KV head are stacked as consecutive rows of one problem, which replaces repeat_kv
// modeling_gemma.py:185-196).
//
// One CTA = QT tiles of 128 query rows (one TMEM lane each) against all keys of its (batch, kv head), key tiles of BN keys:
//   warp 0      TMA producer: Q once (5-D tensor map: dh, group, token, head, batch), then K / V tiles through two rings
//   warp 1      S issuer:    S_j = Q K_j^T  (M=128, N=BN, K=dh; both operands K-major, 128B swizzle), double-buffered in TMEM
//   last warp   P V issuer:  O  += P_j V_j  (M=128, N=dh, K=BN; P K-major from shared memory, V MN-major: the [keys, dh]
//               tile exactly as TMA delivers it).  Two issuing warps because a tcgen05.mma blocks its issuing thread while
//               the tensor pipe's queue is full; both run all 32 lanes with uniform control flow and elect.sync around the
//               instruction, so that descriptors and TMEM addresses live in uniform registers.
//   warps 2..   softmax: thread = query row; S is read from TMEM (tcgen05.ld), P = exp2(s - m) goes to shared memory as
//               bf16 in the swizzled A-operand layout.  The running maximum is only advanced when it grew by more than 2^8
//               (the O accumulator in TMEM then gets rescaled in place), which keeps the accumulator round trip off the
//               common path.  dh <= 128: QT = 2 -- warps 2-5 own query tile 0, warps 6-9 tile 1, sharing every K / V tile;
//               3 of 8 exponentials run on the FMA pipe (exp2_fma), the row sums of dh = 72 come from the tensor core (a
//               ones column in the padded V tile).  dh = 256: QT = 1, four softmax warps (SW = 1).  SW = 2 (eight warps, two
//               column halves of one row exchanging their maxima through shared memory) remains for single-tile problems.
// S_{j+1} is computed while the softmax of tile j runs, and P_j V_j runs while the softmax of tile j+1 runs.
// Zero padding comes from TMA: columns beyond dh (72 -> 80) and key / query rows beyond the sequence are out-of-bounds
// box elements and arrive as zeros; padded keys are masked to -inf before the softmax.
// Timelines of one CTA (clock64 stamps, Params::trace): profiles/tools/attn_prefill_trace.py, profiles/r02h_attn_prefill_trace_*.txt.
#include "common.cuh"
#include "paligemma_b200.h"
#include "tmap.cuh"

#ifndef PG_AP_POLY8
#define PG_AP_POLY8 3
#endif

namespace pg {
namespace ap {

typedef __nv_bfloat16 bf16;

struct Params {
  bf16* o;
  int rows, keys, group, dh;
  long long o_bs, o_ts, o_hs, o_head_off;
  float sl2;  // softmax scale * log2(e)
  const int* key_lens;  // optional [B]: problem b only attends to its first key_lens[b] keys (ragged prompts); else nullptr
  int stagger;          // QT = 2 (tuning): clock cycles the softmax group of query tile 1 idles after its first S tile is ready
  long long* trace;     // optional clock64 stamps of CTA (0,0,0): [role][tile][event], see profiles/tools/attn_prefill_trace.py
};

template <int DH, int QT>
struct Cfg {
  static constexpr int BM = 128;
  static constexpr int BN = (DH > 128 || QT == 2) ? 64 : 128;  // keys per tile
  static constexpr int DHP = (DH + 15) / 16 * 16;       // UMMA K (QK^T) / N (PV) extent: 72 -> 80
  static constexpr int NKB = (DH + 63) / 64;            // 64-column (128 B) boxes along dh
  static constexpr int Q_BOX = BM * 128;                // bytes of one Q box
  static constexpr int KV_BOX = BN * 128;               // bytes of one K / V box
  static constexpr int Q_BYTES = NKB * Q_BOX;           // one query tile
  static constexpr int KV_BYTES = NKB * KV_BOX;
  static constexpr int P_BOX = BM * 128;                // [128 rows x 64 keys]
  static constexpr int P_BYTES = (BN / 64) * P_BOX;
  static constexpr int KST = QT == 2 ? 3 : 2, VST = KST;  // ring depths
  static constexpr int OFF_K = QT * Q_BYTES;
  static constexpr int OFF_V = OFF_K + KST * KV_BYTES;
  static constexpr int OFF_P = OFF_V + VST * KV_BYTES;  // [QT][2 buffers][P_BYTES]
  static constexpr int OFF_BAR = OFF_P + QT * 2 * P_BYTES;
  static constexpr int OFF_X = OFF_BAR + 256;          // float [2 S buffers][2 halves][128 rows]: tile maxima / row sums (SW = 2)
  static constexpr int SMEM = OFF_X + 2048;
  static constexpr int TMEM_COLS = 512;
  static constexpr int COL_S = 0, COL_O = 2 * QT * BN;  // S: [QT][2 buffers][BN] columns, then O: [QT] accumulators
  static constexpr int O_STRIDE = QT == 2 ? 128 : DHP;  // (two accumulators: each starts on a 128-column boundary)
  static constexpr int NBAR = 1 + 2 * KST + 2 * VST + 8 * QT;
  static_assert(2 * QT * BN + (QT - 1) * O_STRIDE + DHP <= 512, "TMEM budget");
  static_assert(8 * NBAR + 8 <= 256, "barrier area");
  static_assert(SMEM <= 232448, "shared memory budget");
};

PG_DEVINL void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
PG_DEVINL void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// registers -> TMEM: lane = TMEM lane, 16 consecutive fp32 columns
PG_DEVINL void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
PG_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory descriptor of an MN-major operand tile as TMA lays it out ([k rows][64 mn columns = 128 B], 128B swizzle):
// 8-row groups along K are 1024 B apart (stride byte offset), 64-column boxes along MN are `box_bytes` apart (leading
// byte offset).  (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.)
PG_DEVINL uint64_t make_sdesc_mn_sw128(uint32_t smem_addr, uint32_t box_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((box_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                  // stride byte offset
  d |= static_cast<uint64_t>(1) << 46;                          // version = 1
  d |= static_cast<uint64_t>(2) << 61;                          // SWIZZLE_128B
  return d;
}


// Waits for two barrier phases with both try_waits in flight together (their latencies overlap instead of adding up).
PG_DEVINL void mbar_wait2(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred PA, PB;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 PA, [%1], %2;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 PB, [%3], %4;\n\t"
      "and.pred PA, PA, PB;\n\t"
      "selp.u32 %0, 1, 0, PA;\n\t}\n"
      : "=r"(ok)
      : "r"(bar_a), "r"(par_a), "r"(bar_b), "r"(par_b)
      : "memory");
  if (ok) return;
  mbar_wait(bar_a, par_a);
  mbar_wait(bar_b, par_b);
}

// One MUFU ex2 per score (ftz; -inf -> +0, so masked keys weigh exactly nothing).
PG_DEVINL float exp2_mufu(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / ALU pipes for a share of the scores (dh <= 128, where the MUFU pipe -- 16 exp2 per clock per SM -- is the
// longest stage of the softmax): round-to-nearest split x = n + f by the 1.5 * 2^23 magic add (no F2I: conversions run on
// the XU pipe too), cubic for 2^f on [-0.5, 0.5] (relative error < 1e-4, more than an order of magnitude below the bf16 rounding of the
// probability that follows), n added into the exponent field.  x <= 8 by the lazy maximum; x = -inf (masked keys) clamps to
// 2^-126, which rounds to a bf16 denormal that weighs nothing.
// (Round 1 measured this as "no gain": at that time the single MMA-issuing thread, not the softmax, paced the kernel.)
PG_DEVINL float exp2_fma(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.05518874f, 0.24261212f);  // minimax cubic of 2^f on [-0.5, 0.5]: max relative error 7.5e-5
  p = fmaf(p, f, 0.69325641f);
  p = fmaf(p, f, 0.99992746f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// QT = query tiles (128 rows each) per CTA.  QT = 2 (dh <= 128): both tiles run against the SAME K / V tiles in shared
// memory -- the kernel is bound by K/V delivery through the crossbar (profiles/r01d_prefill_attn72_ncu_full.csv), and two
// query tiles halve the K/V bytes per FLOP; softmax warps 2-5 own tile 0, warps 6-9 tile 1 (no exchange between them).
template <int DH, int SW, int QT>
__global__ void __launch_bounds__(96 + 128 * SW * QT, 1)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const Params p) {
  static_assert(SW * QT <= 2, "eight softmax warps at most");
  using C = Cfg<DH, QT>;
  constexpr int BM = C::BM, BN = C::BN, DHP = C::DHP, NKB = C::NKB, KST = C::KST, VST = C::VST;
  // dh = 72 is padded to 80 columns: column 72 of every V tile is set to 1.0 in shared memory, so that O[:, 72] accumulates
  // the row sums of the (bf16) probabilities on the tensor core and the softmax warps -- bound by instruction issue, ~4.5
  // instructions per score -- drop the FADD per score.  The lazy rescale of O carries the column along.
  constexpr bool TCSUM = DHP > DH;
  // scores per 8 whose exponential runs on the FMA pipe (exp2_fma): MUFU and FMA / issue time balance near 3 of 8 for dh <= 128;
  // dh = 256 has four times the tensor work per score and is not softmax-bound
  constexpr int POLY8 = DH <= 128 ? PG_AP_POLY8 : 0;
  constexpr uint32_t IDESC_S = make_idesc_bf16(BM, BN);
  constexpr uint32_t IDESC_O = make_idesc_bf16(BM, DHP, 0, 1);  // B (= V) is MN-major
  extern __shared__ __align__(1024) uint8_t smem_ap[];
  const uint32_t sbase = smem_u32(smem_ap);
  if ((sbase & 1023u) != 0) __trap();
  const uint32_t bar0 = sbase + C::OFF_BAR;
  // barriers: q_full, k_full[KST], k_empty[KST], v_full[VST], v_empty[VST], then per query tile s_full[2], s_empty[2],
  // p_full[2] and o_done[2]
  constexpr int S0 = 1 + 2 * KST + 2 * VST;
  const uint32_t q_full = bar0;
  auto k_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bar0 + 8u * (1 + KST + s); };
  auto v_full = [&](int s) { return bar0 + 8u * (1 + 2 * KST + s); };
  auto v_empty = [&](int s) { return bar0 + 8u * (1 + 2 * KST + VST + s); };
  auto s_full = [&](int qt, int s) { return bar0 + 8u * (S0 + qt * 2 + s); };
  auto s_empty = [&](int qt, int s) { return bar0 + 8u * (S0 + 2 * QT + qt * 2 + s); };
  auto p_full = [&](int qt, int s) { return bar0 + 8u * (S0 + 4 * QT + qt * 2 + s); };
  auto o_done = [&](int qt, int s) { return bar0 + 8u * (S0 + 6 * QT + qt * 2 + s); };  // P_j V_j complete: barrier j % 2
  const uint32_t tmem_slot = bar0 + 8u * C::NBAR;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_ap + C::OFF_BAR + 8 * C::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  // profiling stamps (tiles 0..31 of CTA (0,0,0)): role 0 = MMA issuer, 1 = softmax warp 2, 2 = softmax warp 6; 8 events per tile
  const bool tr_on = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
  auto stamp = [&](int role, int j, int ev) {
    if (tr_on && j < 32) p.trace[(role * 32 + j) * 8 + ev] = clock64();
  };
  // ragged batches: keys [key_lens[b], keys) of problem b exist in memory (padding tokens, finite values) but weigh nothing
  const int n_keys = p.key_lens ? max(1, min(p.keys, __ldg(p.key_lens + b))) : p.keys;
  const int n_tiles = (n_keys + BN - 1) / BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KST; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); }
    for (int s = 0; s < VST; ++s) { mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
    for (int qt = 0; qt < QT; ++qt) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(s_full(qt, s), 1); mbar_init(s_empty(qt, s), 4 * SW);
        mbar_init(p_full(qt, s), 4 * SW);
      }
      mbar_init(o_done(qt, 0), 1);
      mbar_init(o_done(qt, 1), 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      mbar_expect_tx(q_full, QT * C::Q_BYTES);
#pragma unroll
      for (int qt = 0; qt < QT; ++qt) {
        const int t0 = ((m_blk * QT + qt) * BM) / p.group;  // first token of this row tile (128 % group == 0)
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb)
          tma_load_5d(sbase + qt * C::Q_BYTES + kb * C::Q_BOX, &tmQ, q_full, kb * 64, 0, t0, h, b);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int ks = j % KST, vs = j % VST;
        mbar_wait(k_empty(ks), ((j / KST) & 1) ^ 1);
        mbar_expect_tx(k_full(ks), C::KV_BYTES);
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb)
          tma_load_4d(sbase + C::OFF_K + ks * C::KV_BYTES + kb * C::KV_BOX, &tmK, k_full(ks), kb * 64, j * BN, h, b);
        mbar_wait(v_empty(vs), ((j / VST) & 1) ^ 1);
        mbar_expect_tx(v_full(vs), C::KV_BYTES);
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb)
          tma_load_4d(sbase + C::OFF_V + vs * C::KV_BYTES + kb * C::KV_BOX, &tmV, v_full(vs), kb * 64, j * BN, h, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================== MMA issuer ================================
    // All 32 lanes run the loop and ONE elected lane issues: with uniform control flow the compiler keeps the descriptors
    // and TMEM addresses in uniform registers.  (Under `if (lane == 0)` every tcgen05.mma paid an ELECT + R2UR.BROADCAST
    // chain per operand, ~115 clk per instruction: the issuer, not the softmax or the tensor pipe, paced the kernel --
    // profiles/r02h_attn_prefill_trace_before.txt.)
    auto issue_s = [&](int j) {  // S_j = Q K_j^T of every query tile into its S buffer j % 2
      const int ks = j % KST, sb = j & 1;
      mbar_wait(k_full(ks), (j / KST) & 1);
      const uint32_t kaddr = sbase + C::OFF_K + ks * C::KV_BYTES;
#pragma unroll
      for (int qt = 0; qt < QT; ++qt) {
        mbar_wait(s_empty(qt, sb), ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + C::COL_S + (qt * 2 + sb) * BN;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DHP / 16; ++k) {
            const uint64_t adesc = make_sdesc_k_sw128(sbase + qt * C::Q_BYTES + (k >> 2) * C::Q_BOX) + 2 * (k & 3);
            const uint64_t bdesc = make_sdesc_k_sw128(kaddr + (k >> 2) * C::KV_BOX) + 2 * (k & 3);
            umma_f16(d_tmem, adesc, bdesc, IDESC_S, k > 0 ? 1u : 0u);
          }
          umma_commit(s_full(qt, sb));
          if (qt == QT - 1) umma_commit(k_empty(ks));
        }
        __syncwarp();
      }
    };
    // The S issuer (this warp) and the P V issuer (the last warp) are separate instruction streams: a tcgen05.mma blocks its
    // issuing thread while the tensor pipe's queue is full, so one stream serialised S_{j+1}, its barrier round trips, the
    // ones column and P_j V_j into 2300 clk per 128-key tile -- longer than the softmax it was supposed to hide behind.
    mbar_wait(q_full, 0);
    for (int j = 0; j < n_tiles; ++j) {
      stamp(0, j, 0);
      issue_s(j);
      stamp(0, j, 1);
    }
  } else if (warp == 2 + 4 * SW * QT) {
    // ============================== P V issuer ================================
    for (int j = 0; j < n_tiles; ++j) {
      const int vs = j % VST, pb = j & 1;
      mbar_wait(v_full(vs), (j / VST) & 1);
      stamp(0, j, 2);
      const uint32_t vaddr = sbase + C::OFF_V + vs * C::KV_BYTES;
      if constexpr (TCSUM) {
        // every lane plants its share of the ones column into the V tile that just landed:
        // column DH (= 72) lives in box DH / 64, 16-byte chunk (DH % 64) / 8, element DH % 8 of key row r (128B swizzle)
        constexpr uint32_t BOX = DH / 64, CH = (DH % 64) / 8, EL = DH % 8;
        for (int r = lane; r < BN; r += 32) {
          const uint32_t addr = vaddr + BOX * C::KV_BOX + r * 128 + ((CH ^ (r & 7)) << 4) + EL * 2;
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<unsigned short>(0x3F80)) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
      }
      stamp(0, j, 3);
#pragma unroll
      for (int qt = 0; qt < QT; ++qt) {
        mbar_wait(p_full(qt, pb), (j >> 1) & 1);
        if (qt == 0) stamp(0, j, 4);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + C::COL_O + qt * C::O_STRIDE;
        const uint32_t paddr = sbase + C::OFF_P + (qt * 2 + pb) * C::P_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BN / 16; ++k) {
            const uint64_t adesc = make_sdesc_k_sw128(paddr + (k >> 2) * C::P_BOX) + 2 * (k & 3);
            const uint64_t bdesc = make_sdesc_mn_sw128(vaddr + k * 2048, C::KV_BOX);  // 16 keys = two 8-row groups
            umma_f16(d_tmem, adesc, bdesc, IDESC_O, (j > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(o_done(qt, pb));
          if (qt == QT - 1) umma_commit(v_empty(vs));
        }
        __syncwarp();
      }
      stamp(0, j, 5);
    }
  } else {
    // ============================== softmax / epilogue ========================
    constexpr int HB = BN / SW;         // score columns of a tile owned by one thread
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int grp = (warp - 2) >> 2;    // second set of four warps: the other column half (SW = 2) or query tile 1 (QT = 2)
    const int half = SW == 2 ? grp : 0;
    const int qt = QT == 2 ? grp : 0;
    const int r = q * 32 + lane;        // row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int row = (m_blk * QT + qt) * BM + r;
    const uint32_t col_s = C::COL_S + qt * 2 * BN, col_o = C::COL_O + qt * C::O_STRIDE;
    float* xch = reinterpret_cast<float*>(smem_ap + C::OFF_X);  // [2][2][128]
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); };  // the two warps of a quadrant
    float m_used = -INFINITY, l_run = 0.f;
    const uint32_t p_row = sbase + C::OFF_P + qt * 2 * C::P_BYTES + r * 128;
    for (int j = 0; j < n_tiles; ++j) {
      const int sb = j & 1;
      const int trole = warp == 2 ? 1 : (warp == 6 ? 2 : 3);
      if (trole < 3) stamp(trole, j, 0);
      // S_j is ready AND P_{j-2} V_{j-2} has completed (its P buffer is the one this tile writes): both barrier round trips
      // (~150 clk each even when the phase completed long ago) are in flight together.  P_{j-1} V_{j-1}, which the old loop
      // waited for at the end of every tile, only matters to the rare O rescale below.
      if (j >= 2) mbar_wait2(s_full(qt, sb), (j >> 1) & 1, o_done(qt, sb), ((j - 2) >> 1) & 1);
      else mbar_wait(s_full(qt, sb), (j >> 1) & 1);
      if constexpr (QT == 2) {
        if (j == 0 && qt == 1 && p.stagger > 0) {
          const long long t0 = clock64();
          while (clock64() - t0 < p.stagger) {}
        }
      }
      if (trole < 3) stamp(trole, j, 1);
      tc_fence_after();
      float s[HB];
      {
        // every 16-column slice is requested before the one wait: the slices' TMEM round trips overlap instead of queueing
        // (ncu at 4096 keys: tensor pipe 28 %, XU 45 %, issue 33 % -- nothing saturated, the softmax warps sit in latencies)
        uint32_t v[HB / 16][16];
#pragma unroll
        for (int c0 = 0; c0 < HB; c0 += 16) tmem_ld16(lane_addr + col_s + sb * BN + half * HB + c0, v[c0 / 16]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < HB; ++i) s[i] = __uint_as_float(v[i / 16][i % 16]);  // raw scores: the scale rides on the FFMA below
      }
      tc_fence_before();
      __syncwarp();
      if (trole < 3) stamp(trole, j, 2);
      if (lane == 0) mbar_arrive(s_empty(qt, sb));  // the S buffer may be overwritten by S_{j+2}
      // The softmax warps are bound by instruction issue (ncu: ~9 instructions per score, XU pipe 39 %, tensor pipe 24 %), so
      // the per-score work is kept to FMNMX, FFMA, MUFU, FADD and half an F2FP: keys are only masked in the one tile that
      // has padding, and the softmax scale is folded into the exponent's FFMA (scale > 0: max commutes with it).
      const int nvalid = n_keys - j * BN - half * HB;  // columns of this thread that are real keys
      if (nvalid < HB) {
#pragma unroll
        for (int i = 0; i < HB; ++i)
          if (i >= nvalid) s[i] = -INFINITY;
      }
      float mx;
      {
        // eight independent running maxima, then a tree: a single FMNMX chain over HB scores is HB dependent instructions
        float m8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m8[i] = s[i];
#pragma unroll
        for (int i = 8; i < HB; ++i) m8[i & 7] = fmaxf(m8[i & 7], s[i]);
        mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
      }
      mx *= p.sl2;
      if constexpr (SW == 2) {  // row maximum over both halves (buffer sb is rewritten two tiles later, one barrier apart)
        xch[(sb * 2 + half) * 128 + r] = mx;
        pair_sync();
        mx = fmaxf(mx, xch[(sb * 2 + (half ^ 1)) * 128 + r]);
      }
      if (trole < 3) stamp(trole, j, 3);
      // lazy running maximum: only move it when it grew by more than 8 (log2 units); probabilities stay <= 2^8
      const bool need = mx > m_used + 8.0f;
      const float m_new = need ? mx : m_used;
      const float alpha = need ? exp2f(m_used - m_new) : 1.0f;  // first tile: m_used = -inf -> 0 (O is overwritten)
      m_used = m_new;
      float sum = 0.f;
      const uint32_t pdst = p_row + sb * C::P_BYTES;
      // Three separate passes over the thread's HB scores (exponent arguments, then the MUFU exponentials back to back, then
      // pack + store): interleaved chunk by chunk, each exp2 -> pack -> store chain exposed the ~40 clk MUFU latency with two
      // or three exponentials in flight (900 clk per 64 scores for a warp ALONE on its sub-partition's MUFU pipe, which
      // could do them in 512 -- profiles/r02h_attn_prefill_trace_qt2.txt).
#pragma unroll
      for (int i = 0; i < HB; ++i) s[i] = fmaf(s[i], p.sl2, -m_new);
      // (an earlier variant with floorf / float->int range reduction was 40 % SLOWER: those conversions run on the XU pipe
      //  as well, so it added XU work instead of removing it)
#pragma unroll
      for (int i = 0; i < HB; ++i) s[i] = (POLY8 > 0 && (i & 7) >= 8 - POLY8) ? exp2_fma(s[i]) : exp2_mufu(s[i]);
#pragma unroll
      for (int cc = 0; cc < HB / 8; ++cc) {  // 16-byte chunks of 8 keys
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = s[cc * 8 + 2 * e], p1 = s[cc * 8 + 2 * e + 1];
          if constexpr (!TCSUM) sum += p0 + p1;
          pk[e] = pack_bf16(p0, p1);
        }
        const int c = half * (HB / 8) + cc;  // chunk index inside the whole tile row
        const uint32_t addr = pdst + (c >> 3) * C::P_BOX + (((c & 7) ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
      }
      l_run = l_run * alpha + sum;  // (SW == 2: the sum over this thread's columns only; the halves are added at the end)
      if (trole < 3) stamp(trole, j, 4);
      if (j > 0) {
        if (__any_sync(0xffffffffu, need)) {  // (both warps of a quadrant see the same maxima, hence the same decision)
          mbar_wait(o_done(qt, (j - 1) & 1), ((j - 1) >> 1) & 1);  // P_{j-1} V_{j-1} has completed: O may be touched
          tc_fence_after();
#pragma unroll 1
          for (int c0 = half * 16; c0 < DHP; c0 += 16 * SW) {  // the halves take alternate 16-column slices of O
            uint32_t v[16];
            tmem_ld16(lane_addr + col_o + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st16(lane_addr + col_o + c0, v);
          }
          tmem_st_wait();
        }
      }
      if (trole < 3) stamp(trole, j, 5);
      fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(qt, sb));
      if (trole < 3) stamp(trole, j, 6);
    }
    // ---- epilogue: O / l -> bf16 ----
    if constexpr (SW == 2 && !TCSUM) {  // total row sum = sum over both halves
      pair_sync();            // the partner has read the last tile's maximum: the exchange area is free
      xch[half * 128 + r] = l_run;
      pair_sync();
      l_run += xch[(half ^ 1) * 128 + r];
    }
    mbar_wait(o_done(qt, (n_tiles - 1) & 1), ((n_tiles - 1) >> 1) & 1);
    tc_fence_after();
    if constexpr (TCSUM) {  // the row sum is column DH of the accumulator (ones column of V); any warp of the quadrant may read it
      uint32_t v[16];
      tmem_ld16(lane_addr + col_o + (DH / 16) * 16, v);
      tmem_ld_wait();
      l_run = __uint_as_float(v[DH % 16]);
    }
    const float inv = 1.0f / l_run;
    const bool row_ok = row < p.rows;
    bf16* orow = p.o + b * p.o_bs + h * p.o_head_off + static_cast<long long>(row / p.group) * p.o_ts +
                 static_cast<long long>(row % p.group) * p.o_hs;
#pragma unroll 1
    for (int c0 = half * 16; c0 < DHP; c0 += 16 * SW) {
      uint32_t v[16];
      tmem_ld16(lane_addr + col_o + c0, v);
      tmem_ld_wait();
      if (row_ok) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(__uint_as_float(v[2 * i]) * inv, __uint_as_float(v[2 * i + 1]) * inv);
        if (c0 + 8 <= p.dh) *reinterpret_cast<uint4*>(orow + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (c0 + 16 <= p.dh) *reinterpret_cast<uint4*>(orow + c0 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// n-D bf16 tensor map, 128B swizzle, OOB -> zeros.  dims/strides innermost first; strides in ELEMENTS (dim 0 is contiguous).
static int make_tmap_nd(CUtensorMap* m, const void* ptr, int rank, const long long* dims, const long long* strides, const int* box) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return PG_ERR_DRIVER;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = static_cast<cuuint64_t>(dims[i]);
    bx[i] = static_cast<cuuint32_t>(box[i]);
    estr[i] = 1;
    if (i > 0) {
      long long sb = strides[i] * 2;
      if (dims[i] == 1 && (sb <= 0 || (sb % 16) != 0)) sb = 16;  // the stride of an extent-1 dimension is never used
      if (sb <= 0 || (sb % 16) != 0) return PG_ERR_ARG;
      gstr[i - 1] = static_cast<cuuint64_t>(sb);
    }
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PG_OK : PG_ERR_TMAP;
}

template <int DH, int SW, int QT>
static int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Params& p, int B, int H, cudaStream_t st) {
  using C = Cfg<DH, QT>;
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(attn_prefill_tc_kernel<DH, SW, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess) {
      cudaGetLastError();
      return PG_ERR_CUDA;
    }
    configured[dev] = true;
  }
  dim3 grid((p.rows + C::BM * QT - 1) / (C::BM * QT), H, B);
  attn_prefill_tc_kernel<DH, SW, QT><<<grid, 96 + 128 * SW * QT, C::SMEM, st>>>(tq, tk, tv, p);
  pg_count_launch(1);
  return cudaGetLastError() == cudaSuccess ? PG_OK : PG_ERR_CUDA;
}

}  // namespace ap
}  // namespace pg

static int g_force_qt = 0, g_stagger = 500, g_force_sw = 0;
static long long* g_trace = nullptr;
extern "C" int pg_debug_set_attn_prefill_trace(long long* device_buffer) { g_trace = device_buffer; return 0; }
extern "C" int pg_debug_set_attn_prefill(int force_qt, int stagger_clk) {  // tuning sweeps (dh <= 128): query tiles per CTA (0 = automatic), stagger
  g_force_qt = force_qt & 3;
  g_force_sw = (force_qt >> 4) & 3;  // bits 4-5: softmax warps per TMEM lane quadrant of the one-tile variant (0 = automatic)
  if (stagger_clk >= 0) g_stagger = stagger_clk;
  return 0;
}

// Returns PG_OK when the problem was launched on the tcgen05 kernel, 1 when the shape / strides are not supported by it
// (the caller then uses the mma.sync kernel), or a negative error.
int pg_attention_prefill_tc(const void* q, const void* k, const void* v, void* o, int B, int H, int rows, int keys, int dh,
                            int group, long long q_bs, long long q_ts, long long q_hs, long long q_head_off, long long kv_bs,
                            long long kv_ts, long long kv_head_off, long long o_bs, long long o_ts, long long o_hs,
                            long long o_head_off, float scale, const int* key_lens, void* stream) {
  using namespace pg;
  if (dh != 64 && dh != 72 && dh != 256) return 1;
  if (!(scale > 0.f)) return 1;  // the kernel takes row maxima before scaling
  if (group <= 0 || (128 % group) != 0 || (rows % group) != 0) return 1;
  auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  auto ok8 = [](long long s) { return s > 0 && (s % 8) == 0; };
  if (!al16(q) || !al16(k) || !al16(v) || !al16(o)) return 1;
  const int tokens = rows / group;
  if (!ok8(q_ts) || !ok8(kv_ts) || !ok8(o_ts) || (group > 1 && (!ok8(q_hs) || !ok8(o_hs)))) return 1;
  if ((H > 1 && (!ok8(q_head_off) || !ok8(kv_head_off) || !ok8(o_head_off))) || (B > 1 && (!ok8(q_bs) || !ok8(kv_bs)))) return 1;
  CUtensorMap tq, tk, tv;
  int rc;
  {
    const long long dims[5] = {dh, group, tokens, H, B};
    const long long strides[5] = {1, q_hs, q_ts, q_head_off, q_bs};
    const int box[5] = {64, group, 128 / group, 1, 1};
    if ((rc = ap::make_tmap_nd(&tq, q, 5, dims, strides, box)) != PG_OK) return rc == PG_ERR_ARG ? 1 : rc;
  }
  // query tiles per CTA (dh <= 128).  Two tiles share every K / V tile and give each SM sub-partition two INDEPENDENT softmax
  // warps (one per query tile) instead of two halves of one row in lock step: with the S and P V issuers split into two
  // warps, 4096 keys run at 1.02 ms vs 1.21 ms for the one-tile variant (profiles/r02h_attn72_variants.txt).
  const int qt = dh > 128 ? 1 : (g_force_qt ? g_force_qt : (rows > 128 ? 2 : 1));
  const int BN = (dh > 128 || qt == 2) ? 64 : 128;
  {
    const long long dims[4] = {dh, keys, H, B};
    const long long strides[4] = {1, kv_ts, kv_head_off, kv_bs};
    const int box[4] = {64, BN, 1, 1};
    if ((rc = ap::make_tmap_nd(&tk, k, 4, dims, strides, box)) != PG_OK) return rc == PG_ERR_ARG ? 1 : rc;
    if ((rc = ap::make_tmap_nd(&tv, v, 4, dims, strides, box)) != PG_OK) return rc == PG_ERR_ARG ? 1 : rc;
  }
  ap::Params p;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.rows = rows; p.keys = keys; p.group = group; p.dh = dh;
  p.o_bs = o_bs; p.o_ts = o_ts; p.o_hs = o_hs; p.o_head_off = o_head_off;
  p.sl2 = scale * 1.4426950408889634f;
  p.key_lens = key_lens;
  p.trace = g_trace;
  p.stagger = g_stagger;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // softmax warps per TMEM lane quadrant: two (column halves) for the one-tile dh <= 128 variant, else one
  const int sw = qt == 2 ? 1 : (g_force_sw ? g_force_sw : (dh > 128 ? 1 : 2));
#define PG_AP_LAUNCH1(DHV)                                                           \
  if (sw == 1) return ap::launch<DHV, 1, 1>(tq, tk, tv, p, B, H, st);             \
  return ap::launch<DHV, 2, 1>(tq, tk, tv, p, B, H, st);
#define PG_AP_LAUNCH2(DHV)                                                           \
  if (qt == 2) return ap::launch<DHV, 1, 2>(tq, tk, tv, p, B, H, st);
  switch (dh) {
    case 64: PG_AP_LAUNCH2(64) PG_AP_LAUNCH1(64)
    case 72: PG_AP_LAUNCH2(72) PG_AP_LAUNCH1(72)
    default: PG_AP_LAUNCH1(256)
  }
#undef PG_AP_LAUNCH1
#undef PG_AP_LAUNCH2
}
