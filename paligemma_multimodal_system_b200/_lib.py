"""ctypes binding of libpaligemma_b200.so (the C ABI declared in include/paligemma_b200.h).

There is no CPU fallback: `lib()` raises if the shared library is missing, and every wrapper raises RuntimeError on a
non-zero return code.  Tensors are passed as raw device pointers (`tensor.data_ptr()`); torch only owns the memory and
the stream.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PG_LIB_PATH") or os.path.join(_HERE, "libpaligemma_b200.so")  # PG_LIB_PATH: A/B builds (profiling)

EPI_BF16, EPI_F32, EPI_ATOMIC_F32, EPI_GEGLU = 0, 1, 2, 3

_ERRORS = {-1: "PG_ERR_ARG", -2: "PG_ERR_CUDA", -3: "PG_ERR_DRIVER", -4: "PG_ERR_TMAP", -5: "PG_ERR_ARCH"}

p, i32, i64, f32, u64 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_ulonglong

# name -> argtypes (restype is int unless noted); must match include/paligemma_b200.h
SIGNATURES = {
    "pg_abi_version": [],
    "pg_check_device": [],
    "pg_num_sms": [],
    "pg_launch_count": [],
    "pg_set_pdl": [i32],
    "pg_prefetch_l2": [p, i64, i32, i32, p],
    "pg_debug_set_gemm_trace": [p],
    "pg_debug_set_gemm_bn": [i32],
    "pg_debug_set_rmsnorm_early_trigger": [i32],
    "pg_debug_set_attn_prefill": [i32, i32],
    "pg_debug_set_attn_prefill_trace": [p],
    "pg_debug_set_gemm_pair": [i32, i32],
    "pg_debug_set_attn_trace": [p],
    "pg_debug_set_sampler_cluster": [i32],
    "pg_debug_topp_retries": [],
    "pg_debug_topp_trace": [p],
    "pg_debug_set_topp_bracket": [i32],
    "pg_gemm_bf16": [p, i64, p, i64, p, i64, p, p, i64, i32, i32, i32, i32, i32, f32, i32, i32, p],
    "pg_gemm_bf16_fused": [p, i64, p, i64, p, i64, p, p, i64, i32, i32, i32, i32, i32, f32, i32, i32, p, p],
    "pg_gemm_qkv_rope": [p, i64, p, i64, p, i64, i32, i32, i32, i32, i32, p, p, p, p, p, p, i32, i32, i32, p],
    "pg_pack_gate_up": [p, p, p, i32, i32, p],
    "pg_cast_f32_bf16": [p, p, i64, p],
    "pg_layernorm": [p, p, p, p, p, i32, i32, f32, p],
    "pg_rmsnorm": [p, p, p, i32, i32, f32, p],
    "pg_resample_h_u8": [p, p, i32, i32, i32, p, p, i32, p],
    "pg_resample_v_u8_norm": [p, p, i32, i32, i32, p, p, i32, p, p],
    "pg_im2col": [p, p, i32, i32, i32, i32, i32, i32, p],
    "pg_attention_prefill": [p, p, p, p, i32, i32, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, f32, p],
    "pg_attention_prefill_varlen": [p, p, p, p, p, i32, i32, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, f32, p],
    "pg_rope_kv_append": [p, i32, p, p, p, p, p, p, p, p, i32, i32, i32, i32, i32, i32, i32, p, p],
    "pg_attention_decode_fused": [p, p, p, p, p, p, p, p, i32, i32, i32, i32, i32, i32, i32, f32, p],
    "pg_kv_gather": [p, p, p, i32, i32, i32, i32, i32, i32, p],
    "pg_merge_embeddings": [p, p, p, p, p, p, p, p, i32, i32, i32, i32, i64, i64, f32, f32, p],
    "pg_merge_scan": [p, p, p, p, p, p, i32, i32, i32, i64, i64, p],
    "pg_merge_text": [p, p, p, p, i32, i32, i32, i32, f32, p],
    "pg_embed_tokens": [p, p, p, p, i32, i32, i32, f32, f32, i64, i64, p],
    "pg_argmax": [p, i64, p, i32, i32, p],
    "pg_sample_top_p": [p, i64, p, p, i32, i32, f32, f32, u64, p, p],
    "pg_argmax_stats": [p, i64, p, i64, p, i32, i32, p],
    "pg_sample_top_p_stats": [p, i64, p, i64, p, i32, i32, f32, f32, u64, p, p, p],
    "pg_advance_decode": [p, p, p, p, i32, p, i32, p],
    "pg_advance_decode_slots": [p, p, i32, p, p, p, p, i32, p],
}
_RESTYPE = {"pg_launch_count": i64}



class GemmFusion(C.Structure):
    """Mirror of PgGemmFusion (include/paligemma_b200.h)."""
    _fields_ = [
        ("zero_buf", p), ("zero_count", i64),
        ("stats", p), ("stats_ld", i64), ("stat_c", f32), ("resid_row_mod", i32), ("out_row_map", p),
    ]


_lib = None


def lib():
    """Loads the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m paligemma_multimodal_system_b200.build` "
                "(there is no CPU / PyTorch fallback for the kernels)")
        l = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, i32)
        _lib = l
        if os.environ.get("PG_PDL", "1") == "0":
            l.pg_set_pdl(0)
        if os.environ.get("PG_GEMM_PAIR", "1") == "0":  # A/B runs: one-CTA prefill GEMM only
            l.pg_debug_set_gemm_pair(0, 0)
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {_ERRORS.get(rc, rc)}")


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def num_sms():
    return int(lib().pg_num_sms())


def require_device():
    """The product path needs an sm_100 GPU and the compiled kernels; anything else is an error, not a fallback."""
    if not torch.cuda.is_available():
        raise RuntimeError("paligemma_multimodal_system_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
    check(lib().pg_check_device(), "pg_check_device")


# ------------------------------------------------------------------------------------------------------------------
# thin typed wrappers (argument checking that needs tensor metadata lives here; kernels check the rest)
# ------------------------------------------------------------------------------------------------------------------
def gemm(x, w, out, *, mode, bias=None, resid=None, act_gelu=False, scale=1.0, swap=-1, split_k=1, features=None, resid_row_mod=0,
         out_row_map=None):
    """out[t,f] (mode-dependent) from x [T,K] bf16 and w [F,K] bf16 (nn.Linear layout).  resid_row_mod = N > 0: `resid` is an
    [N, F] table and token t takes row t % N (position embeddings).  out_row_map int32 [T]: token t goes to row out_row_map[t]
    of `out` (projector rows scattered to their `<image>` positions)."""
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.is_cuda and w.is_cuda
    assert x.dim() == 2 and w.dim() == 2 and x.stride(1) == 1 and w.stride(1) == 1 and x.shape[1] == w.shape[1]
    T, K = x.shape
    F = w.shape[0] if features is None else features
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    if resid is not None:
        assert resid.dtype == torch.float32 and resid.stride(-1) == 1
    if resid_row_mod or out_row_map is not None:
        fu = GemmFusion()
        if resid_row_mod:
            assert resid is not None and resid.shape[0] >= resid_row_mod
            fu.resid_row_mod = int(resid_row_mod)
        if out_row_map is not None:
            assert out_row_map.dtype == torch.int32 and out_row_map.numel() >= T and out_row_map.is_contiguous()
            fu.out_row_map = out_row_map.data_ptr()
        check(lib().pg_gemm_bf16_fused(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0), ptr(bias),
                                       ptr(resid), 0 if resid is None else resid.stride(0), T, F, K, mode, int(act_gelu), float(scale),
                                       swap, split_k, C.addressof(fu), stream()), "pg_gemm_bf16_fused")
        return out
    check(lib().pg_gemm_bf16(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0), ptr(bias),
                             ptr(resid), 0 if resid is None else resid.stride(0), T, F, K, mode, int(act_gelu), float(scale),
                             swap, split_k, stream()), "pg_gemm_bf16")
    return out


LOG2E = 1.4426950408889634


def gemm_fused(x, w, out, *, mode, zero_buf=None, bias=None, split_k=1, stats=None, inv_temperature=1.0):
    """Decode-step (swap-AB, tokens <= 128) GEMM with the chores of pg_gemm_bf16_fused:
      zero_buf   fp32 tensor zero-filled after the dependency wait (split-K accumulator of a later kernel);
      stats      fp32 [nseg, T, 2] (nseg = 4 * ceil(F / 128)): lm_head segment statistics (max, sum exp2) at
                 `inv_temperature`, for pg_sample_top_p_stats / pg_argmax_stats (mode EPI_F32 only)."""
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.stride(1) == 1 and w.stride(1) == 1 and x.shape[1] == w.shape[1]
    T, K = x.shape
    fu = GemmFusion()
    if zero_buf is not None:
        assert zero_buf.dtype == torch.float32 and zero_buf.is_contiguous()
        fu.zero_buf, fu.zero_count = zero_buf.data_ptr(), zero_buf.numel()
    if stats is not None:
        assert stats.dtype == torch.float32 and stats.dim() == 3 and stats.shape[2] == 2 and stats.is_contiguous()
        assert stats.shape[0] >= 4 * ((w.shape[0] + 127) // 128) and stats.shape[1] >= T
        fu.stats, fu.stats_ld, fu.stat_c = stats.data_ptr(), stats.shape[1], float(inv_temperature) * LOG2E
    check(lib().pg_gemm_bf16_fused(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0), ptr(bias), 0, 0, T,
                                   w.shape[0], K, mode, 0, 1.0, 1, split_k, C.addressof(fu), stream()), "pg_gemm_bf16_fused")
    return out


def layernorm(x, gamma, beta, eps, out_bf16=None, out_f32=None):
    rows, D = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    check(lib().pg_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ptr(out_bf16), ptr(out_f32), rows, D, float(eps),
                             stream()), "pg_layernorm")


def rmsnorm(x, w, out_bf16, eps=1e-6):
    rows, D = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    check(lib().pg_rmsnorm(x.data_ptr(), w.data_ptr(), out_bf16.data_ptr(), rows, D, float(eps), stream()), "pg_rmsnorm")


def gemm_residual(x, w, h, bias=None):
    """h += x @ w.T (+ bias): the residual-stream GEMMs (out_proj / fc2 / o_proj / down_proj).  With enough output tiles
    the epilogue adds the fp32 residual in place; with few tiles (small token counts: latency path, decode) the K
    dimension is split over CTAs and the partials are red.add-ed into h so that the whole GPU streams the weights."""
    T, K = x.shape
    F = w.shape[0]
    swap = 1 if T <= 128 else 0
    kb = (K + 63) // 64
    if swap:
        tiles = (F + 127) // 128
    else:
        tiles = ((T + 127) // 128) * ((F + 255) // 256)
        if kb < 128:
            # few-token prefill (latency path), short/medium K: count 128x64 tiles -- the kernel narrows its N tile before
            # the host splits K, because every split costs a full fp32 red.add pass over the output tile (measured at
            # T = 256: fc2 14 splits x 256 columns 34 us -> 4 splits x 64 columns 15.6 us; o_proj 6 splits 25.6 -> 1 split
            # 15.4 us).  Long reductions (down_proj, K = 16384) keep wide tiles + splits: 64-column tiles re-read the
            # activation tile from L2 too often.
            tiles = ((T + 127) // 128) * ((F + 63) // 64)
    if tiles >= 96 or kb < 8 or T > 512:  # (large-batch prefill keeps the deterministic, atomics-free epilogue)
        return gemm(x, w, h, mode=EPI_F32, bias=bias, resid=h, swap=swap)
    split = max(1, min(kb // 4, (2 * 148 if swap else 148) // tiles))
    return gemm(x, w, h, mode=EPI_ATOMIC_F32, bias=bias, swap=swap, split_k=split)
