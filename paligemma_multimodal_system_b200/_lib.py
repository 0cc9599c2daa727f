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
    "pg_launch_count": [],
    "pg_set_pdl": [i32],
    "pg_debug_set_gemm_trace": [p],
    "pg_debug_set_gemm_bn": [i32],
    "pg_debug_set_attn_trace": [p],
    "pg_debug_topp_retries": [],
    "pg_debug_topp_trace": [p],
    "pg_debug_set_topp_bracket": [i32],
    "pg_debug_set_decode_gemm_cta_trace": [p],
    "pg_debug_decode_gemm_blocks_per_sm": [i32],
    "pg_debug_decode_gemm_max_clusters": [i32],
    "pg_debug_set_decode_gemm_trace": [p],
    "pg_gemm_bf16": [p, i64, p, i64, p, i64, p, p, i64, i32, i32, i32, i32, i32, f32, i32, i32, p],
    "pg_gemm_bf16_colnorm": [p, i64, p, i64, p, i64, p, p, i64, i32, i32, i32, i32, i32, f32, i32, i32, p, i32, f32, p],
    "pg_gemm_decode": [p, i64, p, i64, i32, i32, i32, i32, i32, p, i64, p, p, i32, f32, p, i64, p, p, p],
    "pg_decode_prologue": [p, p, p, p, p, p, p, p, i64, i32, i32, i32, f32, f32, i64, i64, p],
    "pg_pack_gate_up": [p, p, p, i32, i32, p],
    "pg_cast_f32_bf16": [p, p, i64, p],
    "pg_layernorm": [p, p, p, p, p, i32, i32, f32, p],
    "pg_rmsnorm": [p, p, p, i32, i32, f32, p, i64, p, i64, p],
    "pg_prefetch_l2": [p, i64, i32, i32, p],
    "pg_resample_h_u8": [p, p, i32, i32, i32, p, p, i32, p],
    "pg_resample_v_u8_norm": [p, p, i32, i32, i32, p, p, i32, p, p],
    "pg_im2col": [p, p, i32, i32, i32, i32, i32, i32, p],
    "pg_add_pos_emb": [p, p, i32, i32, i32, p],
    "pg_attention_prefill": [p, p, p, p, i32, i32, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, f32, p],
    "pg_attention_prefill_varlen": [p, p, p, p, p, i32, i32, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, f32, p],
    "pg_rope_kv_append": [p, i32, p, p, p, p, p, p, p, p, i32, i32, i32, i32, i32, i32, i32, p, p],
    "pg_attention_decode": [p, p, p, p, p, p, p, i32, i32, i32, i32, i32, i32, i32, f32, p],
    "pg_attention_decode_workspace_floats": [i32, i32, i32, i32],
    "pg_attention_decode_fused": [p, p, p, p, p, p, p, p, i32, i32, i32, i32, i32, i32, i32, f32, p],
    "pg_kv_gather": [p, p, p, i32, i32, i32, i32, i32, i32, p],
    "pg_merge_embeddings": [p, p, p, p, p, p, p, p, i32, i32, i32, i32, i64, i64, f32, f32, p],
    "pg_embed_tokens": [p, p, p, p, i32, i32, i32, f32, f32, i64, i64, p],
    "pg_argmax": [p, i64, p, i32, i32, p],
    "pg_sample_top_p": [p, i64, p, p, i32, i32, f32, f32, u64, p, p],
    "pg_advance_decode": [p, p, p, p, i32, p, i32, p],
    "pg_advance_decode_slots": [p, p, i32, p, p, p, p, i32, p],
    "pg_decode_step": [p, p],
    "pg_decode_step_encode_maps": [p, p, p, p, p, i32, i32, i32, i32, i32, i32, i32, i32],
}
_RESTYPE = {"pg_attention_decode_workspace_floats": i64, "pg_launch_count": i64}



class DecodeStepArgs(C.Structure):
    """Mirror of PgDecodeStepArgs (include/paligemma_b200.h)."""
    _fields_ = [
        ("tensor_maps", p),
        ("L", i32), ("B", i32), ("D", i32), ("F", i32), ("Hq", i32), ("Hkv", i32), ("dh", i32), ("V", i32),
        ("split_qkv", i32), ("split_o", i32), ("split_down", i32),
        ("cur_tok", p), ("embed", p), ("img", p), ("n_img", i32), ("text_scale", f32), ("img_scale", f32),
        ("pad_token", i64), ("image_token", i64),
        ("h", p), ("hn", p), ("qkv", p), ("att", p), ("mid", p), ("logits", p),
        ("ln1", p), ("ln2", p), ("norm_w", p), ("head_b", p), ("eps", f32),
        ("k_pages", p), ("v_pages", p), ("layer_stride", i64), ("page_table", p), ("pos", p), ("kv_len", p), ("inv_freq", p),
        ("max_pages", i32), ("page_size", i32), ("scale", f32), ("barrier_state", p), ("trace", p), ("trace_cta", i32),
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
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {_ERRORS.get(rc, rc)}")


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def require_device():
    """The product path needs an sm_100 GPU and the compiled kernels; anything else is an error, not a fallback."""
    if not torch.cuda.is_available():
        raise RuntimeError("paligemma_multimodal_system_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
    check(lib().pg_check_device(), "pg_check_device")


# ------------------------------------------------------------------------------------------------------------------
# thin typed wrappers (argument checking that needs tensor metadata lives here; kernels check the rest)
# ------------------------------------------------------------------------------------------------------------------
def gemm(x, w, out, *, mode, bias=None, resid=None, act_gelu=False, scale=1.0, swap=-1, split_k=1, features=None):
    """out[t,f] (mode-dependent) from x [T,K] bf16 and w [F,K] bf16 (nn.Linear layout)."""
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.is_cuda and w.is_cuda
    assert x.dim() == 2 and w.dim() == 2 and x.stride(1) == 1 and w.stride(1) == 1 and x.shape[1] == w.shape[1]
    T, K = x.shape
    F = w.shape[0] if features is None else features
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    if resid is not None:
        assert resid.dtype == torch.float32 and resid.stride(-1) == 1
    check(lib().pg_gemm_bf16(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0), ptr(bias),
                             ptr(resid), 0 if resid is None else resid.stride(0), T, F, K, mode, int(act_gelu), float(scale),
                             swap, split_k, stream()), "pg_gemm_bf16")
    return out


DEC_F32, DEC_RESID_NORM = 0, 1


def gemm_colnorm(x, w, out, *, mode, ss_in, norm_dim, eps=1e-6, bias=None, swap=1):
    """Swap-AB (decode) GEMM whose epilogue applies the producer's RMSNorm factor rsqrt(ss_in[t]/norm_dim + eps) per token."""
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.stride(1) == 1 and w.stride(1) == 1
    assert ss_in.dtype == torch.float32 and ss_in.numel() >= x.shape[0]
    T, K = x.shape
    check(lib().pg_gemm_bf16_colnorm(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0), ptr(bias),
                                     0, 0, T, w.shape[0], K, mode, 0, 1.0, swap, 1, ss_in.data_ptr(), int(norm_dim), float(eps),
                                     stream()), "pg_gemm_bf16_colnorm")
    return out


def gemm_decode(x, w, out, *, mode, cluster_k, bias=None, ss_in=None, norm_dim=0, eps=1e-6, hb=None, norm_w=None, ss_out=None):
    """Cluster split-K decode GEMM (csrc/gemm_decode.cu): DEC_F32 writes out fp32; DEC_RESID_NORM adds into the fp32 residual
    stream `out` and emits hb = bf16(out * (1 + norm_w)) and ss_out += sum(out^2) for the next fused RMSNorm."""
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.stride(1) == 1 and w.stride(1) == 1
    assert out.dtype == torch.float32 and out.stride(1) == 1
    T, K = x.shape
    assert w.shape[1] == K
    if mode == DEC_RESID_NORM:
        assert hb.dtype == torch.bfloat16 and norm_w.dtype == torch.float32 and ss_out.dtype == torch.float32
    check(lib().pg_gemm_decode(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), T, w.shape[0], K, mode, int(cluster_k),
                               out.data_ptr(), out.stride(0), ptr(bias), ptr(ss_in), int(norm_dim), float(eps), ptr(hb),
                               0 if hb is None else hb.stride(0), ptr(norm_w), ptr(ss_out), stream()), "pg_gemm_decode")
    return out


def layernorm(x, gamma, beta, eps, out_bf16=None, out_f32=None):
    rows, D = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    check(lib().pg_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ptr(out_bf16), ptr(out_f32), rows, D, float(eps),
                             stream()), "pg_layernorm")


def rmsnorm(x, w, out_bf16, eps=1e-6, zero_buf=None, prefetch=None, prefetch_bytes=None):
    rows, D = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    pf_bytes = 0 if prefetch is None else (prefetch.numel() * prefetch.element_size() if prefetch_bytes is None else prefetch_bytes)
    check(lib().pg_rmsnorm(x.data_ptr(), w.data_ptr(), out_bf16.data_ptr(), rows, D, float(eps), ptr(zero_buf),
                           0 if zero_buf is None else zero_buf.numel(), ptr(prefetch), pf_bytes, stream()), "pg_rmsnorm")


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
