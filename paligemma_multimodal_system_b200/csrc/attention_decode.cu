// Decode-step attention, third generation: RoPE + KV append + attention over the paged bf16 cache in one launch
// (modeling_gemma.py:285-339 at q_len == 1 + KVCache.update :18-57), GQA group <= 8.
//
// One thread-block cluster per (sequence, kv head); every rank owns a contiguous range of 64-key pages that it pulls with
// TMA tensor loads into 128B-swizzled shared memory.  The page loads are issued BEFORE griddepcontrol.wait: the cache
// rows of earlier positions, kv_len and the page table were written by kernels that completed before this one could
// start (every kernel of the decode chain waits on its predecessor before it triggers its dependents), only the fp32
// qkv row of the new token comes from the immediately preceding GEMM.
//
// Work split inside a CTA (8 warps), chosen so that NO accumulator ever has to be merged across warps:
//   scores : warp w computes Q K^T for keys [8w, 8w+8) of every resident page (mma.sync m16n8k16, Q fragments live in
//            registers for the whole kernel) and drops the scaled, masked scores into shared memory;
//   softmax: warp r owns query head r: running max / sum across rounds, probabilities written as bf16;
//   P V    : warp w owns the output columns [w*dh/8, (w+1)*dh/8) over ALL keys, so its fp32 accumulators are final.
// Long contexts are processed in rounds of 3 pages with the usual online-softmax rescale between rounds.  The ranks of a
// cluster merge their (unnormalised) partial rows through distributed shared memory.
#include <cooperative_groups.h>

#include "common.cuh"
#include "paligemma_b200.h"
#include "tmap.cuh"

namespace pg {
namespace ad {

typedef __nv_bfloat16 bf16;
namespace cg = cooperative_groups;

struct Params {
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
  long long* trace;
};

template <int DH>
struct Cfg {
  static constexpr int BLOCK_N = 64;
  static constexpr int NBUF = 3;
  static constexpr int NT = 256;
  static constexpr int NBOX = DH / 64;
  static constexpr int BOX_BYTES = BLOCK_N * 128;
  static constexpr int KV_BYTES = NBOX * BOX_BYTES;
  static constexpr int SLOT_BYTES = 2 * KV_BYTES;
  static constexpr int QLD = DH + 8;            // bf16 elements per Q row (+16 B: conflict-free ldmatrix)
  static constexpr int RK = NBUF * BLOCK_N;     // keys per round
  static constexpr int SLD = RK + 4;            // fp32 score row pitch
  static constexpr int PLD = RK + 8;            // bf16 probability row pitch
  static constexpr int RLD = DH + 2;            // partial row: DH accumulators, m (log2 domain), l
  static constexpr int OFF_Q = NBUF * SLOT_BYTES;
  static constexpr int OFF_S = OFF_Q + 16 * QLD * 2;
  static constexpr int OFF_P = OFF_S + 8 * SLD * 4;
  static constexpr int OFF_PART = OFF_P + 16 * PLD * 2;
  static constexpr int SMEM = OFF_PART + 8 * RLD * 4;
  static constexpr int NDP = DH / 16;                    // 16-column output blocks
  static constexpr int DPW = NDP >= 8 ? NDP / 8 : 1;     // blocks per warp
};

template <int DH>
__global__ void __launch_bounds__(256) attn_decode_v3_kernel(const __grid_constant__ CUtensorMap tmK,
                                                             const __grid_constant__ CUtensorMap tmV, const Params p) {
  using C = Cfg<DH>;
  constexpr int BLOCK_N = C::BLOCK_N, NBUF = C::NBUF, NT = C::NT, HALF = DH / 2;
  constexpr int BOX_BYTES = C::BOX_BYTES, KV_BYTES = C::KV_BYTES, SLOT_BYTES = C::SLOT_BYTES;
  constexpr int QLD = C::QLD, SLD = C::SLD, PLD = C::PLD, RLD = C::RLD;
  extern __shared__ __align__(1024) uint8_t smem_ad[];
  uint8_t* ring = smem_ad;
  bf16* Qs = reinterpret_cast<bf16*>(smem_ad + C::OFF_Q);    // [16][QLD]
  float* Ss = reinterpret_cast<float*>(smem_ad + C::OFF_S);  // [8][SLD]
  bf16* Ps = reinterpret_cast<bf16*>(smem_ad + C::OFF_P);    // [16][PLD]
  float* part = reinterpret_cast<float*>(smem_ad + C::OFF_PART);  // [8][RLD]
  __shared__ bf16 new_k[DH], new_v[DH];
  __shared__ __align__(8) uint64_t bars[NBUF];
  __shared__ float s_m[8], s_l[8], s_alpha[8];
  const uint32_t ring_u32 = smem_u32(ring);
  if ((ring_u32 & 1023u) != 0) __trap();

  cg::cluster_group cluster = cg::this_cluster();
  const int CS = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.x / CS;
  const int b = seq / p.Hkv, hk = seq % p.Hkv;
  const int group = p.Hq / p.Hkv;  // <= 8
  // profiling stamps: CTA 0 writes trace[0..7], CTA 1 (its cluster peer) writes trace[256 + 0..7]
  const bool tr = p.trace != nullptr && blockIdx.x < 2 && threadIdx.x == 0;
  long long* const trp = p.trace + (blockIdx.x == 1 ? 256 : 0);
  if (tr) trp[0] = clock64();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < NBUF; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 8) {
    s_m[threadIdx.x] = -INFINITY;
    s_l[threadIdx.x] = 0.f;
    s_alpha[threadIdx.x] = 0.f;
  }
  // probabilities of the padding rows (8..15) and the padding rows of Q stay zero for the whole kernel
  for (int idx = threadIdx.x; idx < 16 * PLD / 2; idx += NT) reinterpret_cast<uint32_t*>(Ps)[idx] = 0u;
  for (int idx = threadIdx.x; idx < 16 * QLD / 2; idx += NT) reinterpret_cast<uint32_t*>(Qs)[idx] = 0u;
  __syncthreads();

  // ---- geometry (all of it produced by kernels that completed before this one started) ----
  const int len = __ldg(p.kv_len + b);
  const int n_tiles = (len + BLOCK_N - 1) / BLOCK_N;
  const int tps = (n_tiles + CS - 1) / CS;
  const int t_begin = min(n_tiles, rank * tps), t_end = min(n_tiles, t_begin + tps);
  const int n_my = t_end - t_begin;
  const int new_slot = len - 1;
  const int new_tile = new_slot / BLOCK_N;
  const bool owns_new = (new_tile >= t_begin && new_tile < t_end);
  const long long kv_ts = static_cast<long long>(p.Hkv) * DH;
  const int* ptab = p.page_table + b * p.max_pages;

  auto load_page = [&](int tile, int slot) {  // one thread
    const int page = __ldg(ptab + tile);
    const uint32_t bar = smem_u32(&bars[slot]);
    const uint32_t dst = ring_u32 + slot * SLOT_BYTES;
    mbar_expect_tx(bar, SLOT_BYTES);
#pragma unroll
    for (int bx = 0; bx < C::NBOX; ++bx) {
      tma_load_2d(dst + bx * BOX_BYTES, &tmK, bar, hk * DH + bx * 64, page * BLOCK_N, kEvictFirst);
      tma_load_2d(dst + KV_BYTES + bx * BOX_BYTES, &tmV, bar, hk * DH + bx * 64, page * BLOCK_N, kEvictFirst);
    }
  };
  if (threadIdx.x == 128) {
    for (int i = 0; i < NBUF && i < n_my; ++i) load_page(t_begin + i, i);
  }
  // byte offset of element (row r, column c) inside a swizzled K (or V) page
  auto swz = [&](int r, int c) -> int { return (c >> 6) * BOX_BYTES + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2; };

  // the page that receives the new token's key / value: known before the wait (the table belongs to earlier kernels), so the
  // owning rank does not pay a second dependent L2 round trip after it
  const int new_page = owns_new ? __ldg(ptab + new_tile) : 0;
  // (measured, not kept: computing sincos before the wait, and signalling the rank merge through per-rank mbarriers
  //  instead of the cluster barrier -- the cluster-scope release of the pushed partials costs the ~1.5 us, not the barrier)
  griddep_wait();  // the qkv row of the new token
  if (tr) trp[1] = clock64();
  if (threadIdx.x == 0) griddep_launch_dependents();

  // ---- RoPE (rotate-half, modeling_gemma.py:138-151) of the query heads (+ the new key), all 256 threads ----
  {
    constexpr int NPART = NT / HALF;  // thread groups that share one frequency index
    const int i = threadIdx.x % HALF, partid = threadIdx.x / HALF;
    const int W = (p.Hq + 2 * p.Hkv) * DH;
    const float* __restrict__ row = p.qkv + static_cast<long long>(b) * W;
    const float posf = static_cast<float>(__ldg(p.pos + b));
    const float freq = __ldg(p.inv_freq + i);
    float x1[8], x2[8], kx1 = 0.f, kx2 = 0.f, vx1 = 0.f, vx2 = 0.f;
    const bool do_k = owns_new && partid == 0, do_v = owns_new && partid == NPART - 1;
#pragma unroll
    for (int gi = 0; gi < 8; ++gi) {
      const int g = partid + gi * NPART;
      if (g < group) {
        const float* qh = row + (hk * group + g) * DH;
        x1[gi] = __ldcg(qh + i);
        x2[gi] = __ldcg(qh + i + HALF);
      }
    }
    if (do_k) {
      const float* kh = row + (p.Hq + hk) * DH;
      kx1 = __ldcg(kh + i); kx2 = __ldcg(kh + i + HALF);
    }
    if (do_v) {
      const float* vh = row + (p.Hq + p.Hkv + hk) * DH;
      vx1 = __ldcg(vh + i); vx2 = __ldcg(vh + i + HALF);
    }
    float sn, cs;
    sincosf(posf * freq, &sn, &cs);
#pragma unroll
    for (int gi = 0; gi < 8; ++gi) {
      const int g = partid + gi * NPART;
      if (g < group) {
        Qs[g * QLD + i] = __float2bfloat16(x1[gi] * cs - x2[gi] * sn);
        Qs[g * QLD + i + HALF] = __float2bfloat16(x2[gi] * cs + x1[gi] * sn);
      }
    }
    if (do_k || do_v) {
      const long long off = (static_cast<long long>(new_page) * BLOCK_N + (new_slot - new_tile * BLOCK_N)) * kv_ts + hk * DH;
      if (do_k) {
        const bf16 k1 = __float2bfloat16(kx1 * cs - kx2 * sn), k2 = __float2bfloat16(kx2 * cs + kx1 * sn);
        p.k_pages[off + i] = k1; p.k_pages[off + i + HALF] = k2;  // KVCache.update (modeling_gemma.py:18-57)
        new_k[i] = k1; new_k[i + HALF] = k2;
      }
      if (do_v) {
        const bf16 v1 = __float2bfloat16(vx1), v2 = __float2bfloat16(vx2);
        p.v_pages[off + i] = v1; p.v_pages[off + i + HALF] = v2;
        new_v[i] = v1; new_v[i + HALF] = v2;
      }
    }
  }
  __syncthreads();  // Q, new_k / new_v staged
  if (tr) trp[2] = clock64();

  // Q fragments for every k-step stay in registers (the query does not change across pages)
  uint32_t qf[DH / 16][4];
  {
    const uint32_t q_addr = smem_u32(Qs + (lane & 15) * QLD + (lane >> 4) * 8);
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) ldmatrix_x4(q_addr + ks * 32, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  }
  float o[2 * C::DPW][4];
#pragma unroll
  for (int i = 0; i < 2 * C::DPW; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  const bool pv_active = warp * C::DPW < C::NDP;
  const int row = lane >> 2;  // query head of this lane's accumulator rows (rows +8 are padding)

  for (int r0 = 0; r0 < n_my; r0 += NBUF) {
    const int npg = min(NBUF, n_my - r0);
    const uint32_t parity = (r0 / NBUF) & 1;
    for (int i = 0; i < npg; ++i) mbar_wait(smem_u32(&bars[i]), parity);
    if (tr && r0 == 0) trp[3] = clock64();
    if (owns_new && new_tile >= t_begin + r0 && new_tile < t_begin + r0 + npg) {
      // the TMA has landed: overwrite the new token's (stale) row with the freshly rotated key / value
      uint8_t* Kb = ring + (new_tile - t_begin - r0) * SLOT_BYTES;
      const int r = new_slot - new_tile * BLOCK_N;
      for (int k = threadIdx.x; k < DH; k += NT) {
        *reinterpret_cast<bf16*>(Kb + swz(r, k)) = new_k[k];
        *reinterpret_cast<bf16*>(Kb + KV_BYTES + swz(r, k)) = new_v[k];
      }
    }
    __syncthreads();

    // ---- scores: warp w -> keys [8w, 8w+8) of every page of the round ----
    for (int pg = 0; pg < npg; ++pg) {
      const uint32_t k_base = ring_u32 + pg * SLOT_BYTES;
      const int kr = warp * 8 + (lane & 7);
      const uint32_t k_row = k_base + kr * 128;
      float se[4] = {0.f, 0.f, 0.f, 0.f}, so[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < DH / 32; ++j) {
        const int chunk = 4 * j + (lane >> 3);  // 16-byte chunk along dh
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(k_row + (chunk >> 3) * BOX_BYTES + (((chunk & 7) ^ (kr & 7)) << 4), b0, b1, b2, b3);
        mma_bf16_16816(se, qf[2 * j], b0, b1);
        mma_bf16_16816(so, qf[2 * j + 1], b2, b3);
      }
      const int key0 = (t_begin + r0 + pg) * BLOCK_N + warp * 8 + (lane & 3) * 2;
      float v0 = (se[0] + so[0]) * p.sl2, v1 = (se[1] + so[1]) * p.sl2;
      if (key0 >= len) v0 = -INFINITY;
      if (key0 + 1 >= len) v1 = -INFINITY;
      if (row < group) *reinterpret_cast<float2*>(Ss + row * SLD + pg * BLOCK_N + warp * 8 + (lane & 3) * 2) = make_float2(v0, v1);
    }
    __syncthreads();

    // ---- softmax statistics: warp r <-> query head r ----
    if (warp < group) {
      const int nk = npg * BLOCK_N;
      const float* srow = Ss + warp * SLD;
      float sv[C::RK / 32];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < C::RK / 32; ++i) {
        const int k = lane + 32 * i;
        sv[i] = k < nk ? srow[k] : -INFINITY;
        mx = fmaxf(mx, sv[i]);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_old = s_m[warp];
      const float m_new = fmaxf(m_old, mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      float sum = 0.f;
      bf16* prow = Ps + warp * PLD;
#pragma unroll
      for (int i = 0; i < C::RK / 32; ++i) {
        const int k = lane + 32 * i;
        const float pr = exp2f(sv[i] - m_safe);
        sum += pr;
        if (k < nk) prow[k] = __float2bfloat16(pr);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      if (lane == 0) {
        const float alpha = exp2f(m_old - m_safe);
        s_alpha[warp] = alpha;
        s_m[warp] = m_new;
        s_l[warp] = s_l[warp] * alpha + sum;
      }
    }
    __syncthreads();

    // ---- P V: warp w -> output columns [w*dh/8, (w+1)*dh/8) over all keys of the round ----
    if (pv_active) {
      if (r0 > 0) {
        const float a = s_alpha[row & 7];
#pragma unroll
        for (int i = 0; i < 2 * C::DPW; ++i) { o[i][0] *= a; o[i][1] *= a; }
      }
      const uint32_t p_addr = smem_u32(Ps + (lane & 15) * PLD + (lane >> 4) * 8);
      for (int pg = 0; pg < npg; ++pg) {
        const uint32_t v_base = ring_u32 + pg * SLOT_BYTES + KV_BYTES;
#pragma unroll
        for (int kk = 0; kk < BLOCK_N / 16; ++kk) {
          uint32_t a[4];
          ldmatrix_x4(p_addr + (pg * BLOCK_N + kk * 16) * 2, a[0], a[1], a[2], a[3]);
          const int vr = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const uint32_t v_row = v_base + vr * 128;
          const int v_sub = lane >> 4;
#pragma unroll
          for (int d = 0; d < C::DPW; ++d) {
            const int dp = warp * C::DPW + d;
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(v_row + (dp >> 2) * BOX_BYTES + (((((dp & 3) << 1) + v_sub) ^ (vr & 7)) << 4), b0, b1, b2, b3);
            mma_bf16_16816(o[2 * d], a, b0, b1);
            mma_bf16_16816(o[2 * d + 1], a, b2, b3);
          }
        }
      }
    }
    __syncthreads();  // the round's pages, scores and probabilities are consumed
    if (r0 + NBUF < n_my && threadIdx.x == 128) {
      fence_proxy_async_smem();
      for (int i = 0; i < NBUF && r0 + NBUF + i < n_my; ++i) load_page(t_begin + r0 + NBUF + i, i);
    }
  }
  if (tr) trp[4] = clock64();

  // ---- result of this CTA: rows < group, this warp's columns ----
  const long long hq0 = static_cast<long long>(b) * p.Hq + hk * group;
  if (CS == 1) {
    if (pv_active && row < group) {
      const float inv = 1.f / s_l[row];
#pragma unroll
      for (int i = 0; i < 2 * C::DPW; ++i) {
        const int col = (warp * C::DPW + (i >> 1)) * 16 + (i & 1) * 8 + (lane & 3) * 2;
        *reinterpret_cast<uint32_t*>(p.out + (hq0 + row) * DH + col) = pack_bf16(o[i][0] * inv, o[i][1] * inv);
      }
    }
    return;
  }

  // ---- merge the ranks: rank q finalises the columns [q*DH/CS, (q+1)*DH/CS).  Every rank PUSHES its unnormalised partial
  //      columns (and its row max / sum) into the owners' shared memory, one cluster barrier later all reads are local,
  //      so no rank has to stay alive for its peers. ----
  __shared__ float s_ml[8][8][2];  // [source rank][row][max (log2 domain), sum]
  __shared__ float s_rw[8][8];     // weight (incl. 1 / row sum) of rank q's partial for row r
  const int cols_per = DH / CS;
  float* recv = part;              // [CS source ranks][8 rows][cols_per]
  if (pv_active && row < group) {
#pragma unroll
    for (int i = 0; i < 2 * C::DPW; ++i) {
      const int col = (warp * C::DPW + (i >> 1)) * 16 + (i & 1) * 8 + (lane & 3) * 2;
      const int q = col / cols_per;
      float* dst = cluster.map_shared_rank(recv, q) + (rank * 8 + row) * cols_per + (col - q * cols_per);
      *reinterpret_cast<float2*>(dst) = (n_my > 0) ? make_float2(o[i][0], o[i][1]) : make_float2(0.f, 0.f);
    }
  }
  if (threadIdx.x < group * CS) {
    const int r = threadIdx.x % group, q = threadIdx.x / group;
    float* dst = &cluster.map_shared_rank(&s_ml[0][0][0], q)[(rank * 8 + r) * 2];
    dst[0] = s_m[r];  // (-inf, 0) when this rank had no keys
    dst[1] = s_l[r];
  }
  cluster.sync();
  if (tr) trp[5] = clock64();
  if (threadIdx.x < group) {
    const int r = threadIdx.x;
    float M = -INFINITY;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < CS) M = fmaxf(M, s_ml[q][r][0]);
    float Lsum = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < CS) Lsum += s_ml[q][r][1] * exp2f(s_ml[q][r][0] - M);
    const float inv = 1.f / Lsum;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < CS) s_rw[r][q] = exp2f(s_ml[q][r][0] - M) * inv;
  }
  __syncthreads();
  {
    const int c_lo = rank * cols_per;
    const int ncol2 = cols_per / 2;
    for (int idx = threadIdx.x; idx < group * ncol2; idx += NT) {
      const int r = idx / ncol2, lc = 2 * (idx % ncol2);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q < CS) {
          const float2 v = *reinterpret_cast<const float2*>(recv + (q * 8 + r) * cols_per + lc);
          a0 += v.x * s_rw[r][q];
          a1 += v.y * s_rw[r][q];
        }
      }
      *reinterpret_cast<uint32_t*>(p.out + (hq0 + r) * DH + c_lo + lc) = pack_bf16(a0, a1);
    }
  }
  if (tr) trp[6] = clock64();
}

template <int DH>
static int launch(const Params& p, int num_pages, int cluster_size, cudaStream_t st) {
  using C = Cfg<DH>;
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(attn_decode_v3_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess) {
      cudaGetLastError();
      return PG_ERR_CUDA;
    }
    configured[dev] = true;
  }
  CUtensorMap tmK, tmV;
  int rc;
  const long long cols = static_cast<long long>(p.Hkv) * DH;
  if ((rc = make_tmap_2d(&tmK, p.k_pages, static_cast<long long>(num_pages) * 64, cols, cols, 64)) != PG_OK) return rc;
  if ((rc = make_tmap_2d(&tmV, p.v_pages, static_cast<long long>(num_pages) * 64, cols, cols, 64)) != PG_OK) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(p.B * p.Hkv * cluster_size));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = C::SMEM;
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
  if (cudaLaunchKernelEx(&cfg, attn_decode_v3_kernel<DH>, tmK, tmV, p) != cudaSuccess) {
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  return PG_OK;
}

}  // namespace ad
}  // namespace pg

// called by pg_attention_decode_fused (attention.cu) for GQA groups <= 8
int pg_attention_decode_v3(const float* qkv, const int* pos, const int* kv_len, const float* inv_freq, void* k_pages,
                           void* v_pages, const int* page_table, void* out, int B, int Hq, int Hkv, int dh, int num_pages,
                           int max_pages, float sl2, int cluster_size, long long* trace, void* stream) {
  pg::ad::Params p;
  p.qkv = qkv; p.pos = pos; p.kv_len = kv_len; p.inv_freq = inv_freq;
  p.k_pages = static_cast<__nv_bfloat16*>(k_pages); p.v_pages = static_cast<__nv_bfloat16*>(v_pages);
  p.page_table = page_table; p.out = static_cast<__nv_bfloat16*>(out);
  p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.max_pages = max_pages; p.sl2 = sl2; p.trace = trace;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dh) {
    case 64: return pg::ad::launch<64>(p, num_pages, cluster_size, st);
    case 256: return pg::ad::launch<256>(p, num_pages, cluster_size, st);
    default: return PG_ERR_ARG;
  }
}
