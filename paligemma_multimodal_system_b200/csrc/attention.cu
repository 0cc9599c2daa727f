// C-ABI entry points of the attention kernels (bf16 in/out, fp32 online softmax); the kernels live in
//   attention_prefill_tc.cu : tcgen05 / TMEM flash attention for the SigLIP tower (modeling_siglip.py:96-136, dh = 72) and the
//                             Gemma prefill (modeling_gemma.py:307-339 with the all-zero mask of modeling_paligemma.py:154-156,
//                             dh = 256; the Hq query heads of one KV head are stacked as consecutive rows, so repeat_kv
//                             (modeling_gemma.py:185-196) never materialises)
//   attention_decode.cu     : q_len = 1 decode over the paged bf16 KV cache, fused with RoPE and the cache append.
// There is exactly one kernel per entry point: a shape the kernel cannot express is an argument error, never a fallback.
#include "common.cuh"
#include "paligemma_b200.h"

int pg_attention_prefill_tc(const void* q, const void* k, const void* v, void* o, int B, int H, int rows, int keys, int dh,
                            int group, long long q_bs, long long q_ts, long long q_hs, long long q_head_off, long long kv_bs,
                            long long kv_ts, long long kv_head_off, long long o_bs, long long o_ts, long long o_hs,
                            long long o_head_off, float scale, const int* key_lens, void* stream);  // attention_prefill_tc.cu

int pg_attention_decode_v3(const float* qkv, const int* pos, const int* kv_len, const float* inv_freq, void* k_pages,
                           void* v_pages, const int* page_table, void* out, int B, int Hq, int Hkv, int dh, int num_pages,
                           int max_pages, float sl2, int cluster_size, long long* trace, void* stream);  // attention_decode.cu

extern "C" int pg_attention_prefill(const void* q, const void* k, const void* v, void* o, int B, int H, int rows, int keys,
                                    int dh, int group, long long q_bs, long long q_ts, long long q_hs, long long q_head_off,
                                    long long kv_bs, long long kv_ts, long long kv_head_off, long long o_bs, long long o_ts,
                                    long long o_hs, long long o_head_off, float scale, void* stream) {
  if (B <= 0 || H <= 0 || rows <= 0 || keys <= 0 || group <= 0) return PG_ERR_ARG;
  if (B > 65535 || H > 65535) return PG_ERR_ARG;
  // (pg_attention_prefill_tc returns 1 for a shape / stride set its TMA maps cannot express)
  const int rc = pg_attention_prefill_tc(q, k, v, o, B, H, rows, keys, dh, group, q_bs, q_ts, q_hs, q_head_off, kv_bs, kv_ts,
                                         kv_head_off, o_bs, o_ts, o_hs, o_head_off, scale, nullptr, stream);
  return rc > 0 ? PG_ERR_ARG : rc;
}

extern "C" int pg_attention_prefill_varlen(const void* q, const void* k, const void* v, void* o, const int* key_lens, int B, int H,
                                           int rows, int keys, int dh, int group, long long q_bs, long long q_ts, long long q_hs,
                                           long long q_head_off, long long kv_bs, long long kv_ts, long long kv_head_off,
                                           long long o_bs, long long o_ts, long long o_hs, long long o_head_off, float scale,
                                           void* stream) {
  if (B <= 0 || H <= 0 || rows <= 0 || keys <= 0 || group <= 0 || key_lens == nullptr) return PG_ERR_ARG;
  if (B > 65535 || H > 65535) return PG_ERR_ARG;
  const int rc = pg_attention_prefill_tc(q, k, v, o, B, H, rows, keys, dh, group, q_bs, q_ts, q_hs, q_head_off, kv_bs, kv_ts,
                                         kv_head_off, o_bs, o_ts, o_hs, o_head_off, scale, key_lens, stream);
  return rc > 0 ? PG_ERR_ARG : rc;
}

static long long* g_attn_trace = nullptr;
static int g_attn_trace_idx = 0;
extern "C" int pg_debug_set_attn_trace(long long* p) { g_attn_trace = p; g_attn_trace_idx = 0; return 0; }

extern "C" int pg_attention_decode_fused(const float* qkv, const int* pos, const int* kv_len, const float* inv_freq,
                                         void* k_pages, void* v_pages, const int* page_table, void* out, int B, int Hq,
                                         int Hkv, int dh, int page_size, int num_pages, int max_pages, float scale,
                                         void* stream) {
  if (B <= 0 || Hq <= 0 || Hkv <= 0 || Hq % Hkv != 0 || Hq / Hkv > 8 || page_size != 64 || max_pages <= 0 || num_pages <= 0)
    return PG_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(k_pages) & 15) || (reinterpret_cast<uintptr_t>(v_pages) & 15)) return PG_ERR_ARG;
  long long* trace = g_attn_trace ? g_attn_trace + 8 * (g_attn_trace_idx++ % 32) : nullptr;
  // cluster size: as many CTAs as fit in ONE wave (the 3-deep page ring allows one CTA per SM), at most one page per
  // rank, at most 8 (portable cluster limit)
  int cs = pg::num_sms() / (B * Hkv);
  if (cs > max_pages) cs = max_pages;
  cs = cs >= 8 ? 8 : cs >= 4 ? 4 : cs >= 2 ? 2 : 1;
  return pg_attention_decode_v3(qkv, pos, kv_len, inv_freq, k_pages, v_pages, page_table, out, B, Hq, Hkv, dh, num_pages,
                                max_pages, scale * 1.4426950408889634f, cs, trace, stream);
}
