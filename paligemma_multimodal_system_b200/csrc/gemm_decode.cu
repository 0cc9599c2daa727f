// Decode-step (q_len == 1, <= 128 tokens) weight-streaming GEMM with an in-cluster split-K reduction.
//
//   acc[f, t] = sum_k W[f, k] * X[t, k]          W: nn.Linear weight [features, K] bf16, X: activations [tokens, K] bf16
//
// Replaces the q/k/v, o_proj and down_proj call sites of the reference at q_len == 1 (modeling_gemma.py:274-278,356,
// 210-218) together with the GemmaRMSNorm that FOLLOWS o_proj / down_proj (modeling_gemma.py:172-181,393-417).
//
// Why a second GEMM kernel: at 64 tokens these matrices have only 16-20 output tiles, so the K dimension has to be split
// over CTAs to keep every SM streaming weights.  gemm_tcgen05.cu reduces the splits with fp32 red.add into global memory
// (2.4 M atomics per down_proj launch; the L2 atomic units make that a 3 us serial tail of every launch, and nobody
// knows the final value, so the RMSNorm needs its own kernel).  Here the S CTAs that share one 128-row output tile form
// a thread-block cluster: every rank keeps its fp32 partial tile in TMEM, scatters it through distributed shared memory
// (rank r receives the token columns [r*BN/S, (r+1)*BN/S) of all partials), and the owner finishes the element:
//
//   mode PG_DEC_F32        out[t, f]  = acc * rs_in[t] + bias[f]
//   mode PG_DEC_RESID_NORM h[t, f]   += acc;   hb[t, f] = bf16(h[t, f] * (1 + norm_w[f]));   ss_out[t] += h[t, f]^2
//
// rs_in[t] = rsqrt(ss_in[t] / norm_dim + eps) is the RMSNorm factor of the PRODUCER of X: the producer stored
// X = bf16(h * (1 + w)) and the per-token sum of squares, and the consumer applies the per-token scalar to the fp32
// accumulator (a per-column scale commutes with the GEMM), so no standalone RMSNorm kernel runs during decode.
//
// Pipeline per CTA (192 threads): warp 0 = TMA producer (the weight slabs are requested BEFORE griddepcontrol.wait, they
// never depend on the previous kernel), warp 1 = tcgen05.mma issuer + TMEM owner, warps 2-5 = epilogue.
#include "common.cuh"
#include "paligemma_b200.h"
#include "tmap.cuh"

namespace pg {
namespace dk {

typedef __nv_bfloat16 bf16;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BUDGET = 112 * 1024;  // two CTAs per SM

__host__ __device__ constexpr int stage_bytes(int BN) { return A_BYTES + BN * BK * 2; }
__host__ __device__ constexpr int red_bytes(int BN) { return BM * BN * 4; }
__host__ __device__ constexpr int num_stages(int BN) {
  int s = (SMEM_BUDGET - 256) / stage_bytes(BN);
  return s > 8 ? 8 : s;
}
__host__ __device__ constexpr int smem_bytes(int BN) {
  int ring = num_stages(BN) * stage_bytes(BN);
  int red = red_bytes(BN);
  return (ring > red ? ring : red) + 256;
}

struct Args {
  int tokens, features, K;
  int mode;
  float* out;  // PG_DEC_F32: fp32 [tokens, ldo];  PG_DEC_RESID_NORM: the fp32 residual stream h (read and written)
  long long ldo;
  const float* bias;
  const float* ss_in;
  float inv_norm_dim, eps;
  bf16* hb;
  long long ldh;
  const float* norm_w;
  float* ss_out;
  long long* trace;  // optional profiling stamps (clock64) written by CTA 0
  long long* cta_trace;  // optional [grid][3]: smid, globaltimer at entry, globaltimer at exit of every CTA
};

PG_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
PG_DEVINL void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
PG_DEVINL void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
PG_DEVINL uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
template <int V>
PG_DEVINL void st_cluster(uint32_t addr, const float* v) {
  if constexpr (V == 4) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
  } else if constexpr (V == 2) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v[0]), "f"(v[1]) : "memory");
  } else {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v[0]) : "memory");
  }
}

template <int BN, int S>
__global__ void __launch_bounds__(NUM_THREADS, 2)
gemm_decode_cluster_kernel(const __grid_constant__ CUtensorMap tmapW, const __grid_constant__ CUtensorMap tmapX, const Args args) {
  constexpr int STAGES = num_stages(BN);
  constexpr int STAGE_BYTES = stage_bytes(BN);
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t IDESC = make_idesc_bf16(BM, BN);
  constexpr int CPR = BN / S;                  // token columns owned by each cluster rank
  constexpr int V = CPR >= 4 ? 4 : CPR;        // floats per remote store
  static_assert(BN % S == 0 && CPR >= 1, "cluster size must divide the token tile");
  static_assert(STAGES >= 3, "pipeline too shallow");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  constexpr int MAIN_BYTES = smem_bytes(BN) - 256;
  const uint32_t bar_base = smem_base + MAIN_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + MAIN_BYTES + 8 * (2 * STAGES + 1));
  float* red = reinterpret_cast<float*>(smem_raw);  // [S][128][CPR] fp32, aliases the (drained) operand ring

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int m_blk = blockIdx.x / S;
  const int total_kb = (args.K + BK - 1) / BK;
  const int kb_per = (total_kb + S - 1) / S;
  const int kb0 = min(total_kb, rank * kb_per);
  const int kb1 = min(total_kb, kb0 + kb_per);
  const int nkb = kb1 - kb0;
  const bool tr = args.trace != nullptr && blockIdx.x == 0;
  if (tr && threadIdx.x == 0) args.trace[0] = clock64();
  if (args.cta_trace != nullptr && threadIdx.x == 0) {
    unsigned smid;
    unsigned long long gt;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    args.cta_trace[3 * blockIdx.x] = smid;
    args.cta_trace[3 * blockIdx.x + 1] = static_cast<long long>(gt);
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmapW);
    tma_prefetch_desc(&tmapX);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
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
    if (lane == 0) {
      const int pre = min(STAGES, nkb);
      for (int s = 0; s < pre; ++s) {
        mbar_expect_tx(full_bar(s), STAGE_BYTES);
        tma_load_2d(smem_base + s * STAGE_BYTES, &tmapW, full_bar(s), (kb0 + s) * BK, m_blk * BM, kEvictFirst);
      }
      if (tr) args.trace[1] = clock64();
      griddep_wait();
      if (tr) args.trace[2] = clock64();
      griddep_launch_dependents();
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        if (i >= pre) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_2d(sa, &tmapW, full_bar(stage), (kb0 + i) * BK, m_blk * BM, kEvictFirst);
        }
        tma_load_2d(sa + A_BYTES, &tmapX, full_bar(stage), (kb0 + i) * BK, 0, kEvictLast);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        const uint64_t adesc = make_sdesc_k_sw128(sa);
        const uint64_t bdesc = make_sdesc_k_sw128(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, IDESC, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(empty_bar(stage));
        if (i == nkb - 1) umma_commit(tfull_bar);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (nkb == 0) mbar_arrive(tfull_bar);  // empty K range: this rank contributes zeros
    }
    __syncwarp();
  } else {
    griddep_wait();  // the residual stream, ss_in and the activations behind the X tensor map come from the previous kernel
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    if (tr && threadIdx.x == 64) args.trace[3] = clock64();
  }

  // ---- cluster barrier 1: every rank's MMAs have completed, i.e. every rank's operand ring is free to be overwritten
  //      by the partial tiles of its peers (and every peer CTA is known to be resident)
  __syncwarp();
  cluster_arrive();
  cluster_wait();
  if (tr && threadIdx.x == 64) args.trace[4] = clock64();

  if (warp >= 2) {
    const int q = warp & 3;
    const int rl = q * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t my_slot = smem_base + static_cast<uint32_t>((rank * BM + rl) * CPR) * 4u;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      if (nkb > 0) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 16; i += V) {
        const int col = c0 + i;
        const int dst_rank = col / CPR;
        const int off = col % CPR;
        st_cluster<V>(mapa(my_slot + off * 4u, dst_rank), v + i);
      }
    }
    tc_fence_before();
    if (tr && threadIdx.x == 64) args.trace[5] = clock64();
  }

  // ---- cluster barrier 2: all partials have landed in their owners' shared memory
  __syncwarp();
  cluster_arrive();
  cluster_wait();
  if (tr && threadIdx.x == 64) args.trace[6] = clock64();

  if (warp >= 2) {
    const int q = warp & 3;
    const int rl = q * 32 + lane;
    const int f = m_blk * BM + rl;
    const bool f_ok = f < args.features;
    float acc[CPR];
#pragma unroll
    for (int c = 0; c < CPR; ++c) acc[c] = 0.f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float* src = red + (s * BM + rl) * CPR;
#pragma unroll
      for (int c = 0; c < CPR; c += V) {
        if constexpr (V == 4) {
          const float4 x = *reinterpret_cast<const float4*>(src + c);
          acc[c] += x.x; acc[c + 1] += x.y; acc[c + 2] += x.z; acc[c + 3] += x.w;
        } else if constexpr (V == 2) {
          const float2 x = *reinterpret_cast<const float2*>(src + c);
          acc[c] += x.x; acc[c + 1] += x.y;
        } else {
          acc[c] += src[c];
        }
      }
    }
    if (args.mode == PG_DEC_F32) {
      const float bias = (args.bias != nullptr && f_ok) ? __ldg(args.bias + f) : 0.f;
#pragma unroll
      for (int c = 0; c < CPR; ++c) {
        const int t = rank * CPR + c;
        if (t < args.tokens && f_ok) {
          float x = acc[c];
          if (args.ss_in != nullptr) x *= rsqrtf(args.ss_in[t] * args.inv_norm_dim + args.eps);
          args.out[static_cast<long long>(t) * args.ldo + f] = x + bias;
        }
      }
    } else {
      const float nw = f_ok ? 1.0f + __ldg(args.norm_w + f) : 0.f;
      // all residual loads first (independent, one L2 round trip), then the stores, then the per-token reductions
      float hv[CPR];
#pragma unroll
      for (int c = 0; c < CPR; ++c) {
        const int t = rank * CPR + c;
        hv[c] = (f_ok && t < args.tokens) ? __ldcg(args.out + static_cast<long long>(t) * args.ldo + f) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < CPR; ++c) {
        const int t = rank * CPR + c;
        const bool ok = f_ok && t < args.tokens;
        const float x = ok ? acc[c] + hv[c] : 0.f;
        if (ok) {
          args.out[static_cast<long long>(t) * args.ldo + f] = x;
          args.hb[static_cast<long long>(t) * args.ldh + f] = __float2bfloat16(x * nw);
        }
        hv[c] = x * x;
      }
#pragma unroll
      for (int c = 0; c < CPR; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hv[c] += __shfl_xor_sync(0xffffffffu, hv[c], o);
      }
#pragma unroll
      for (int c = 0; c < CPR; ++c) {
        const int t = rank * CPR + c;
        if (lane == (c & 31) && t < args.tokens) atomicAdd(args.ss_out + t, hv[c]);
      }
    }
  }

  if (tr && threadIdx.x == 64) args.trace[7] = clock64();
  __syncthreads();
  if (args.cta_trace != nullptr && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    args.cta_trace[3 * blockIdx.x + 2] = static_cast<long long>(gt);
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static long long* g_trace = nullptr;
static long long* g_cta_trace = nullptr;
static int g_trace_idx = 0;

template <int BN, int S>
static int launch(const CUtensorMap& tw, const CUtensorMap& tx, const Args& a, cudaStream_t st) {
  static bool configured = false;
  constexpr int smem = smem_bytes(BN);
  auto kern = gemm_decode_cluster_kernel<BN, S>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();
      return PG_ERR_CUDA;
    }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (S > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      return PG_ERR_CUDA;
    }
    configured = true;
  }
  const int m_blocks = (a.features + BM - 1) / BM;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(m_blocks * S));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pg_pdl_enabled() ? 2 : 1;
  pg_count_launch(1);
  if (cudaLaunchKernelEx(&cfg, kern, tw, tx, a) != cudaSuccess) {
    cudaGetLastError();
    return PG_ERR_CUDA;
  }
  return PG_OK;
}

template <int BN>
static int dispatch_s(int S, const CUtensorMap& tw, const CUtensorMap& tx, const Args& a, cudaStream_t st) {
  switch (S) {
    case 1: return launch<BN, 1>(tw, tx, a, st);
    case 2: return launch<BN, 2>(tw, tx, a, st);
    case 4: return launch<BN, 4>(tw, tx, a, st);
    case 8: return launch<BN, 8>(tw, tx, a, st);
    case 16: return launch<BN, 16>(tw, tx, a, st);
    default: return PG_ERR_ARG;
  }
}

}  // namespace dk
}  // namespace pg

using namespace pg;

template <int BN, int S>
static int max_clusters() {
  auto kern = dk::gemm_decode_cluster_kernel<BN, S>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dk::smem_bytes(BN));
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (S > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(S * 64);
  cfg.blockDim = dim3(dk::NUM_THREADS);
  cfg.dynamicSmemBytes = dk::smem_bytes(BN);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = -1;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
  int per_sm = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, dk::NUM_THREADS, dk::smem_bytes(BN));
  return n * 100 + per_sm;
}
extern "C" int pg_debug_decode_gemm_blocks_per_sm(int smem) {
  auto kern = dk::gemm_decode_cluster_kernel<64, 8>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int per_sm = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, dk::NUM_THREADS, smem);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  int dev = 0, smem_sm = 0, regs_sm = 0, blk_sm = 0, rsv = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
  cudaDeviceGetAttribute(&blk_sm, cudaDevAttrMaxBlocksPerMultiprocessor, dev);
  cudaDeviceGetAttribute(&rsv, cudaDevAttrReservedSharedMemoryPerBlock, dev);
  printf("smem %d -> blocks/SM %d | static smem %zu regs %d maxThreads %d | SM: smem %d regs %d blocks %d reserved/block %d\n", smem, per_sm,
         fa.sharedSizeBytes, fa.numRegs, fa.maxThreadsPerBlock, smem_sm, regs_sm, blk_sm, rsv);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dk::smem_bytes(64));
  return per_sm;
}

/* Profiling aid: how many clusters of `cluster_k` CTAs of the 64-token kernel can be resident at once. */
extern "C" int pg_debug_decode_gemm_max_clusters(int cluster_k) {
  switch (cluster_k) {
    case 2: return max_clusters<64, 2>();
    case 4: return max_clusters<64, 4>();
    case 8: return max_clusters<64, 8>();
    case 16: return max_clusters<64, 16>();
    default: return max_clusters<64, 1>();
  }
}

extern "C" int pg_debug_set_decode_gemm_cta_trace(long long* p) { dk::g_cta_trace = p; return 0; }
extern "C" int pg_debug_set_decode_gemm_trace(long long* p) { dk::g_trace = p; dk::g_trace_idx = 0; return 0; }

extern "C" int pg_gemm_decode(const void* x, long long ldx, const void* w, long long ldw, int tokens, int features, int K,
                              int mode, int cluster_k, float* out, long long ldo, const float* bias, const float* ss_in,
                              int norm_dim, float eps, void* hb, long long ldh, const float* norm_w, float* ss_out,
                              void* stream) {
  if (tokens <= 0 || tokens > 128 || features <= 0 || K <= 0) return PG_ERR_ARG;
  if ((K % 8) != 0 || (ldx % 8) != 0 || (ldw % 8) != 0) return PG_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) return PG_ERR_ARG;
  if (mode != PG_DEC_F32 && mode != PG_DEC_RESID_NORM) return PG_ERR_ARG;
  if (out == nullptr) return PG_ERR_ARG;
  if (mode == PG_DEC_RESID_NORM && (hb == nullptr || norm_w == nullptr || ss_out == nullptr)) return PG_ERR_ARG;
  if (ss_in != nullptr && norm_dim <= 0) return PG_ERR_ARG;
  const int BN = tokens <= 16 ? 16 : tokens <= 32 ? 32 : tokens <= 64 ? 64 : 128;
  const int total_kb = (K + dk::BK - 1) / dk::BK;
  int S = cluster_k;
  if (S != 1 && S != 2 && S != 4 && S != 8 && S != 16) return PG_ERR_ARG;
  while (S > 1 && (S > total_kb || S > BN)) S >>= 1;
  dk::Args a;
  a.tokens = tokens; a.features = features; a.K = K; a.mode = mode; a.out = out; a.ldo = ldo; a.bias = bias;
  a.ss_in = ss_in; a.inv_norm_dim = norm_dim > 0 ? 1.0f / static_cast<float>(norm_dim) : 0.f; a.eps = eps;
  a.hb = static_cast<__nv_bfloat16*>(hb); a.ldh = ldh; a.norm_w = norm_w; a.ss_out = ss_out;
  a.cta_trace = dk::g_cta_trace;
  a.trace = dk::g_trace ? dk::g_trace + 8 * (dk::g_trace_idx++ % 64) : nullptr;
  CUtensorMap tw, tx;
  int rc;
  if ((rc = make_tmap_2d(&tw, w, features, K, ldw, dk::BM)) != PG_OK) return rc;
  if ((rc = make_tmap_2d(&tx, x, tokens, K, ldx, BN)) != PG_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (BN) {
    case 16: return dk::dispatch_s<16>(S, tw, tx, a, st);
    case 32: return dk::dispatch_s<32>(S, tw, tx, a, st);
    case 64: return dk::dispatch_s<64>(S, tw, tx, a, st);
    default: return dk::dispatch_s<128>(S, tw, tx, a, st);
  }
}
