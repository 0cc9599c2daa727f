/*
 * paligemma_b200.h -- C ABI of libpaligemma_b200.so: the sm_100a kernels behind the PaliGemma inference hot path.
 *
 * The reference (prtk1729/Paligemma-MultiModal-System) has no FFI: its boundary is the Python module API
 * (modeling_paligemma.py:257-307 forward, modeling_gemma.py:8-64 KVCache, inference.py:29-106 loop).  The host side of
 * this repo keeps that API in Python (paligemma_multimodal_system_b200/*.py) and reaches the GPU only through the
 * entry points below (ctypes).  Every function:
 *   - takes raw DEVICE pointers, sizes and a cudaStream_t (as void*); no torch types;
 *   - is asynchronous on `stream`; returns PG_OK or a negative PG_ERR_* code (never throws, never falls back to CPU);
 *   - cites the reference call site(s) it replaces.
 * Tensors are row-major; "ld" arguments are row pitches in ELEMENTS.
 */
#ifndef PALIGEMMA_B200_H
#define PALIGEMMA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK 0
#define PG_ERR_ARG (-1)    /* bad shape / alignment / enum */
#define PG_ERR_CUDA (-2)   /* launch or attribute failure (cudaGetLastError) */
#define PG_ERR_DRIVER (-3) /* cuTensorMapEncodeTiled entry point not found */
#define PG_ERR_TMAP (-4)   /* tensor-map encode rejected the geometry */
#define PG_ERR_ARCH (-5)   /* device is not sm_100 */

/* GEMM epilogues */
#define PG_EPI_BF16 0       /* out_bf16[t,f] = act((acc + bias[f]) * scale)                */
#define PG_EPI_F32 1        /* out_f32[t,f]  = (acc + bias[f]) * scale + resid_f32[t,f]    */
#define PG_EPI_ATOMIC_F32 2 /* out_f32[t,f] += (acc + bias[f]) * scale   (split-K, red.add) */
#define PG_EPI_GEGLU 3      /* out_bf16[t,g] = gelu_tanh(gate[t,g]) * up[t,g]; weight rows packed [64 gate | 64 up] */

/* Library / device info.  Returns the ABI version (>0) or PG_ERR_ARCH when the current device is not sm_100. */
int pg_abi_version(void);
int pg_check_device(void);
/* Streaming multiprocessors of the current device (the host side sizes split-K factors and persistent grids with it). */
int pg_num_sms(void);
/* Number of kernel launches this library has issued in the calling process (host-side counter). */
long long pg_launch_count(void);
/* Programmatic dependent launch for the decode-step kernels (default on): the next kernel's prologue and weight
 * prefetch overlap the running kernel; every such kernel executes griddepcontrol.wait before dependent accesses. */
int pg_set_pdl(int on);
/* Profiling aid: when non-NULL, CTA 0 of launch i writes 6 clock64 stamps to buffer[8*(i%64) ..] (device). */
int pg_debug_set_gemm_trace(long long* device_buffer);
int pg_debug_set_rmsnorm_early_trigger(int on); /* A/B runs: the decode RMSNorm triggers its dependents before / after its own dependency wait */
int pg_debug_set_gemm_pair(int mode, int min_tiles); /* A/B runs: mode bit 0 = CTA-pair (cta_group::2) prefill GEMM allowed, bit 1 = banded tile raster OFF, bits 2-3 = epilogue warps of the pair kernel (1 = eight, 2 = four, 0 = automatic) (default mode 1); min_tiles > 0 sets the pair kernel's tile-count threshold */
int pg_debug_set_attn_prefill_trace(long long* device_buffer); /* profiling aid: clock64 stamps of CTA (0,0,0) of the tcgen05 prefill attention, 3 roles x 32 tiles x 8 events */
int pg_debug_set_attn_prefill(int force_qt, int stagger_clk); /* tuning sweeps of the prefill attention: force_qt bits 0-1 = query tiles per CTA for dh <= 128, bits 4-5 = softmax warps per TMEM lane quadrant of the one-tile variant (0 = automatic); stagger_clk >= 0 = start offset of the second softmax group */
int pg_debug_set_gemm_bn(int bn); /* tuning sweeps: force the N tile (64 / 128 / 256) of the token-major GEMM, 0 = automatic */
int pg_debug_set_topp_bracket(int half_width_bins); /* > 0 overrides the estimated-bracket half width of pg_sample_top_p */
int pg_debug_topp_trace(long long* host_out16); /* clock64 stamps of CTA 0 after each phase of the last pg_sample_top_p */
int pg_debug_topp_retries(void); /* rows whose estimated top-p bracket failed verification (second full pass), cumulative */
int pg_debug_set_sampler_cluster(int ctas_per_row); /* tuning sweeps: cluster size of pg_sample_top_p_stats (1/2/4/8), 0 = automatic */
int pg_debug_set_sampler_cluster(int ctas_per_row); /* tuning sweeps: cluster size of pg_sample_top_p_stats (1/2/4/8), 0 = automatic */
int pg_debug_set_attn_trace(long long* device_buffer); /* same for pg_attention_decode_fused (8 stamps per launch) */

/*
 * acc[t,f] = sum_k x[t,k] * w[f,k]  (bf16 in, fp32 accumulate on tcgen05 tensor cores, accumulator in TMEM).
 * Replaces every nn.Linear on the path: modeling_siglip.py:59-62,71-75,156,177-185; modeling_paligemma.py:57,64;
 * modeling_gemma.py:205-218,255-259,274-278,356,484,523; and nn.Conv2d (modeling_siglip.py:258-263) after pg_im2col.
 *   swap:   0 = tokens on the UMMA M axis (prefill), 1 = features on the M axis (decode, tokens <= 128), -1 = auto
 *   split_k: >1 only with PG_EPI_ATOMIC_F32
 */
int pg_gemm_bf16(const void* x, long long ldx, const void* w, long long ldw, void* out, long long ldo,
                 const float* bias, const float* resid, long long ldr, int tokens, int features, int K, int mode,
                 int act_gelu, float scale, int swap, int split_k, void* stream);

/*
 * pg_gemm_bf16 with per-call-site extras; `fusion` may be NULL (= pg_gemm_bf16).  zero_buf and stats belong to the swap-AB
 * (tokens <= 128, decode) kernels, resid_row_mod to the token-major (prefill) kernels.
 *
 *  - zero_buf / zero_count: zero-filled (fp32, count % 4 == 0, 16-byte aligned) after the dependency wait; o_proj uses it to
 *    reset the split-K accumulator of the q/k/v projection once the attention kernel has consumed it.
 *  - stats (PG_EPI_F32 only, no resid / split-K): the lm_head of a decode step (modeling_gemma.py:523-525) also emits, for
 *    every token t and every 32-row vocabulary segment g (g = f / 32), stats[g * stats_ld + t] = (m, s) with
 *    m = max logit of the segment and s = sum exp2((logit - m) * stat_c), stat_c = inv_temperature * log2(e) -- temperature
 *    scaling and the max / partition-function passes of softmax + top-p (inference.py:63-66,90-106) folded into the GEMM
 *    epilogue; consumed by pg_sample_top_p_stats / pg_argmax_stats.  float2 [4 * ceil(features / 128)][stats_ld >= tokens].
 *  - resid_row_mod > 0 (token-major kernels, PG_EPI_F32 with resid): the residual row of token t is t % resid_row_mod, i.e.
 *    resid is a [resid_row_mod, features] table broadcast over the batch -- the position-embedding add of the SigLIP patch
 *    embedding (modeling_siglip.py:289-298) folded into the patch GEMM's epilogue.
 *  - out_row_map (token-major kernels, PG_EPI_F32 without resid): token t is written to output row out_row_map[t] -- the
 *    multimodal projector (modeling_paligemma.py:57-65) scattering its rows, times `scale`, straight to the `<image>` positions of
 *    the merged embedding sequence (masked_scatter of _merge_input_ids_with_image_features, :201-251); map from pg_merge_scan.
 */
typedef struct PgGemmFusion {
  float* zero_buf;
  long long zero_count;
  void* stats;
  long long stats_ld;
  float stat_c;
  int resid_row_mod;
  const int* out_row_map;
} PgGemmFusion;
int pg_gemm_bf16_fused(const void* x, long long ldx, const void* w, long long ldw, void* out, long long ldo,
                       const float* bias, const float* resid, long long ldr, int tokens, int features, int K, int mode,
                       int act_gelu, float scale, int swap, int split_k, const PgGemmFusion* fusion, void* stream);

/*
 * Prefill q/k/v projection with RoPE and the KV-cache append in the GEMM epilogue (modeling_gemma.py:274-302 + KVCache.update
 * :18-57): qkv_out[t, :] = bf16 of [rope(q heads) | rope(k heads) | v heads] of x[t, :] @ w^T, w = [q_proj; k_proj; v_proj]
 * ([(Hq + 2 Hkv) * dh, K] bf16), rotate-half with angle pos[t] * inv_freq[i]; k / v rows are also written into their cache
 * pages (page = page_table[b * max_pages + slot / 64], slot = slot_base[b] + (t - b * tokens_per_seq); k_pages == NULL: no
 * cache).  One head per UMMA N tile: dh in {64, 256}.  The prefill attention reads q / k / v straight out of qkv_out through
 * strided tensor maps, so no separate RoPE / append launch and no second copy of the projections exists.
 */
int pg_gemm_qkv_rope(const void* x, long long ldx, const void* w, long long ldw, void* qkv_out, long long ldo, int tokens, int K,
                     int Hq, int Hkv, int dh, const int* pos, const float* inv_freq, void* k_pages, void* v_pages,
                     const int* page_table, const int* slot_base, int tokens_per_seq, int page_size, int max_pages, void* stream);

/* Packs gate_proj / up_proj [F,K] bf16 into the [64 gate | 64 up] row-interleaved [2F,K] layout PG_EPI_GEGLU expects.
 * (modeling_gemma.py:205-206 weights; F % 64 == 0) */
int pg_pack_gate_up(const void* gate, const void* up, void* packed, int F, int K, void* stream);

/* fp32 -> bf16 cast (weight packing, pixel staging). */
int pg_cast_f32_bf16(const float* src, void* dst, long long n, void* stream);

/* nn.LayerNorm (modeling_siglip.py:199-204,310,319): x fp32 [rows, D] -> y bf16 (and/or fp32 if y_f32 != NULL). */
int pg_layernorm(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, int rows, int D,
                 float eps, void* stream);

/* GemmaRMSNorm (modeling_gemma.py:157-182): y = x * rsqrt(mean(x^2) + eps) * (1 + w); x fp32 -> y bf16.  (Prefill, and the
 * decode norms.) */
int pg_rmsnorm(const float* x, const float* w, void* y_bf16, int rows, int D, float eps, void* stream);

/*
 * Image path of PaliGemmaProcessor (processing_paligemma.py:13-73) on the GPU, bit-exact with the reference's CPU path:
 * Pillow's 8-bit bicubic resize as two separable fixed-point passes, then rescale / normalise / HWC->CHW.
 *   pg_resample_h_u8:      src uint8 [H, W, 3] -> dst uint8 [H, S, 3]     (image.resize, horizontal pass, :17-19)
 *   pg_resample_v_u8_norm: src uint8 [H, Wd, 3] -> out fp32 [3, S, Wd]    (vertical pass + lut[256] = the reference's
 *                          (u * 1/255).astype(float32), (x - 0.5) / 0.5 of :21-33, + the transpose of :71)
 * kk int32 [S, ksize] 22-bit fixed-point weights and bounds int32 [S, 2] = (first input index, count) are built by the
 * host exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc do (image_preprocess.py).
 */
int pg_resample_h_u8(const void* src, void* dst, int H, int W, int S, const int* kk, const int* bounds, int ksize, void* stream);
int pg_resample_v_u8_norm(const void* src, float* out, int H, int Wd, int S, const int* kk, const int* bounds, int ksize,
                          const float* lut, void* stream);

/* SiglipVisionEmbeddings im2col (modeling_siglip.py:258-263,285-297): pixel fp32 [B,C,H,W] -> patches bf16
 * [B*(H/P)*(W/P), Kpad], column order (c, py, px) = Conv2d weight order, zero padded to Kpad. */
int pg_im2col(const float* pixels, void* patches, int B, int C, int H, int W, int P, int Kpad, void* stream);

/*
 * Non-causal softmax(Q K^T * scale) V, flash style (modeling_siglip.py:96-136; modeling_gemma.py:307-339 prefill
 * with the all-zero mask of modeling_paligemma.py:154-156).  bf16 in/out, fp32 softmax.
 * Row r of batch b, head h:  q + b*q_bs + (r / group)*q_ts + (r % group)*q_hs + h*q_head_off  (same for o);
 * K/V rows: k + b*kv_bs + n*kv_ts + h*kv_head_off.  `group` > 1 stacks the query heads of one KV head as consecutive
 * rows (MQA/GQA, replaces repeat_kv modeling_gemma.py:185-196).  All strides in elements.  dh in {64, 72, 256}.
 */
int pg_attention_prefill(const void* q, const void* k, const void* v, void* o, int B, int H, int rows, int keys, int dh,
                         int group, long long q_bs, long long q_ts, long long q_hs, long long q_head_off,
                         long long kv_bs, long long kv_ts, long long kv_head_off, long long o_bs, long long o_ts,
                         long long o_hs, long long o_head_off, float scale, void* stream);

/*
 * pg_attention_prefill for a ragged batch (prompts of different lengths in one prefill; the reference
 * runs B = 1, inference.py:69, and never masks padding, modeling_paligemma.py:154-156): problem b attends only to its first
 * key_lens[b] keys (int32 [B] on the device, clamped to [1, keys]).  Rows past a problem's length are computed and must be
 * ignored by the caller; K/V rows in [key_lens[b], keys) must hold finite values.  tcgen05 kernel only (dh 64 / 72 / 256,
 * 128 % group == 0, 16-byte aligned strides): other shapes return PG_ERR_ARG.
 */
int pg_attention_prefill_varlen(const void* q, const void* k, const void* v, void* o, const int* key_lens, int B, int H,
                                int rows, int keys, int dh, int group, long long q_bs, long long q_ts, long long q_hs,
                                long long q_head_off, long long kv_bs, long long kv_ts, long long kv_head_off,
                                long long o_bs, long long o_ts, long long o_hs, long long o_head_off, float scale,
                                void* stream);

/*
 * RoPE + KV append (modeling_gemma.py:116-151,285-302 and KVCache.update :18-57): reads qkv [T, (Hq+2Hkv)*dh]
 * (bf16, or fp32 when qkv_is_f32), rotates q and k (rotate-half, fp32 angle = pos[t] * inv_freq[i], inv_freq fp32 [dh/2] built by the host exactly as
 * GemmaRotaryEmbedding.__init__ does), writes q_out bf16
 * [T, Hq*dh], optional k_out/v_out bf16 [T, Hkv*dh] (dense, for prefill attention) and appends k/v into the paged
 * cache: page = page_table[b*max_pages + slot/page_size], slot = slot_base[b] + (t - b*tokens_per_seq).
 */
int pg_rope_kv_append(const void* qkv, int qkv_is_f32, const int* pos, void* q_out, void* k_out, void* v_out,
                      void* k_pages, void* v_pages, const int* page_table, const int* slot_base, int B,
                      int tokens_per_seq, int Hq, int Hkv, int dh, int page_size, int max_pages,
                      const float* inv_freq, void* stream);

/*
 * Decode-step attention in ONE launch (modeling_gemma.py:285-339 at q_len == 1 + KVCache.update :18-57): rotates q and
 * the new k (fp32 qkv [B, (Hq+2Hkv)*dh] straight from the split-K QKV GEMM), appends k/v to the paged cache at slot
 * kv_len[b]-1 and attends over kv_len[b] keys.  One thread-block cluster per (sequence, kv head): each rank streams a
 * contiguous range of 64-key pages (TMA tensor loads into 128B-swizzled shared memory) and the ranks merge through
 * distributed shared memory.  num_pages = pages in the pool (k_pages / v_pages are [num_pages, 64, Hkv*dh] bf16).
 * GQA group Hq/Hkv <= 8, dh in {64, 256}.
 */
int pg_attention_decode_fused(const float* qkv, const int* pos, const int* kv_len, const float* inv_freq, void* k_pages,
                              void* v_pages, const int* page_table, void* out, int B, int Hq, int Hkv, int dh,
                              int page_size, int num_pages, int max_pages, float scale, void* stream);

/* Gathers the dense K or V of one layer, [B, Hkv, len, dh] bf16, from the paged cache (KVCache.k_cache / v_cache
 * views, modeling_gemma.py:8-64). */
int pg_kv_gather(const void* pages, const int* page_table, void* dense, int B, int len, int Hkv, int dh, int page_size,
                 int max_pages, void* stream);

/*
 * Embedding gather + image-feature merge (+ position ids)
 * (modeling_paligemma.py:93-128,195,288 and the *sqrt(hidden) of modeling_gemma.py:510-511):
 *   text token : h[b,s,:] = embed[id,:] * text_scale      image token (j-th in row b): h = img[b,j,:] * img_scale
 *   pad token  : 0                                         pos[b,s] = cumsum(mask)[s], 1 where mask == 0
 * embed bf16 [V,D]; img fp32 [B,N,D]; h fp32 [B,S,D]; src_scratch int32 [B,S].  err_flag[0] is set to 1 when a row
 * does not hold exactly N image tokens (the reference leaves this unchecked, modeling_paligemma.py:120-121).
 */
int pg_merge_embeddings(const long long* input_ids, const long long* attn_mask, const void* embed, const float* img,
                        float* h, int* pos, int* src_scratch, int* err_flag, int B, int S, int D, int N,
                        long long image_token, long long pad_token, float text_scale, float img_scale, void* stream);

/*
 * The same merge as two halves around the projector GEMM, so that the image features never take a round trip of their own:
 *   pg_merge_scan  -> pos[b,s], src_scratch[b,s] (j >= 0: j-th image token | -1 text | -2 pad), err_flag, and (dst_row != NULL)
 *                     dst_row[b*N + j] = b*S + s: the merged row of image feature j, the out_row_map of the projector GEMM;
 *   pg_merge_text  -> fills the text / pad rows of h (embedding gather * text_scale, zeros) and leaves the image rows alone.
 */
int pg_merge_scan(const long long* input_ids, const long long* attn_mask, int* pos, int* src_scratch, int* dst_row, int* err_flag,
                  int B, int S, int N, long long image_token, long long pad_token, void* stream);
int pg_merge_text(const long long* input_ids, const int* src_scratch, const void* embed, float* h, int B, int S, int D, int N,
                  float text_scale, void* stream);

/* Decode-step embedding with the same merge rules at q_len = 1; tokens are int32 on the device; img may be NULL. */
int pg_embed_tokens(const int* tokens, const void* embed, const float* img, float* h, int B, int D, int N,
                    float text_scale, float img_scale, long long pad_token, long long image_token, void* stream);

/* Greedy argmax over fp32 logits [B, V] (inference.py:68); ties -> lowest index.  out int32 [B]. */
int pg_argmax(const float* logits, long long ld, int* out, int B, int V, void* stream);

/*
 * Top-p sampling (inference.py:65,90-106) from fp32 logits [B, V]: probs = softmax(logits * inv_temperature);
 * keep every token whose strictly-greater probability mass is <= top_p (= the reference's exclusive-cumsum rule; tied
 * probabilities are kept or dropped together); sample from the renormalised kept set by inverse CDF in vocabulary
 * order with a counter-based RNG keyed by (seed, *step_ptr, row).  out int32 [B]; kept_count optional int32 [B].
 */
int pg_sample_top_p(const float* logits, long long ld, int* out, int* kept_count, int B, int V, float inv_temperature,
                    float top_p, unsigned long long seed, const int* step_ptr, void* stream);

/*
 * The same two samplers fed by the segment statistics the lm_head epilogue emits (PgGemmFusion.stats: per 32-token vocabulary
 * segment, max logit and sum exp2((x - max) * inv_temperature * log2 e)): row maximum, partition function and the segment
 * masses of the inverse-CDF draw come from 8 bytes per segment instead of passes over the fp32 row; only the top-p
 * verification pass (mass of strictly more probable tokens <= top_p, inference.py:96-100) still reads the logits.
 * stats: float2 [ceil(V / 32) or more segment rows][stats_ld >= B]; the statistics must have been produced with the SAME
 * inv_temperature.  seed_ptr (device, optional) overrides `seed`, so that a captured CUDA graph serves every seed.
 * A row whose 64 candidates are all rejected (top_p far below the largest probability) yields its most probable token.
 */
int pg_argmax_stats(const float* logits, long long ld, const void* stats, long long stats_ld, int* out, int B, int V,
                    void* stream);
int pg_sample_top_p_stats(const float* logits, long long ld, const void* stats, long long stats_ld, int* out, int B, int V,
                          float inv_temperature, float top_p, unsigned long long seed, const unsigned long long* seed_ptr,
                          const int* step_ptr, void* stream);

/* After sampling (device-side loop state, CUDA-graph friendly): tok_hist[step*B + b] = next[b]; cur_tok[b] = next[b];
 * counters[c*B + b] += 1 for c < n_counters (position ids, KV write slots, KV lengths); step += 1.  B <= 1024. */
int pg_advance_decode(const int* next, int* tok_hist, int* cur_tok, int* counters, int n_counters, int* step, int B,
                      void* stream);

/* pg_advance_decode for continuous batching (the inference.py:45-79 loop state of B independent requests, one per slot):
 * tok_ring[(step % ring)*B + b] = next[b]; cur_tok[b] = next[b]; the three counters of slot b (int32 [3, B]: position id, KV
 * write slot, KV length) advance only while counters[2][b] < kv_limit[b] -- a slot that has used its budget (finished or
 * idle) is frozen until the host re-arms it; step += 1. */
int pg_advance_decode_slots(const int* next, int* tok_ring, int ring, int* cur_tok, int* counters, const int* kv_limit,
                            int* step, int B, void* stream);

/* Decode chain, no reference counterpart (the reference's nn.Linear weights sit in host caches): asks the bulk-copy engine to
 * pull the contiguous range [ptr, ptr + bytes) -- the weights of a LATER GEMM of the decode step (modeling_gemma.py:205-218
 * gate/up/down projections) -- into L2 and returns at once.  Meant for a forked branch of the step's CUDA graph, while the
 * attention block leaves the HBM idle.  `ctas` (0 = one per SM) paces the stream: every CTA keeps one SM's bulk-copy engine busy
 * (~80 GB/s), so a small grid is a background stream that leaves HBM bandwidth to the critical-path loads.  evict_last != 0
 * tags the lines L2::evict_last (measured: no benefit for the gate/up weights, kept for experiments).  ptr 16-byte aligned,
 * bytes >= 16 (rounded down to a multiple of 16).  Never changes results. */
int pg_prefetch_l2(const void* ptr, long long bytes, int ctas, int evict_last, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PALIGEMMA_B200_H */
