"""tcgen05 GEMM (pg_gemm_bf16) against a plain PyTorch fp32 reference of the same op, through the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w):
    return x.float() @ w.float().t()


def _mk(T, F, K, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(T, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(F, K, device="cuda", generator=g) * 0.05).bfloat16()
    return x, w


def _close(a, b, tol, what):
    err = (a.float() - b.float()).abs().max().item()
    ref = b.float().abs().max().item()
    assert err <= tol * max(ref, 1e-6), f"{what}: max-abs err {err:.4g} vs ref absmax {ref:.4g}"


SHAPES = [
    # (tokens, features, K)
    (128, 256, 64), (128, 256, 256), (260, 2560, 2048), (300, 1281, 256), (1024, 4304, 1152), (512, 1152, 4304),
    (256, 1152, 640), (130, 64, 128),
]


@pytest.mark.parametrize("T,F,K", SHAPES)
def test_gemm_prefill_bf16_out(T, F, K):
    from paligemma_multimodal_system_b200 import _lib
    x, w = _mk(T, F, K)
    bias = torch.randn(F, device="cuda")
    out = torch.full((T, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.gemm(x, w, out, mode=_lib.EPI_BF16, bias=bias, swap=0)
    torch.cuda.synchronize()
    _close(out, _ref(x, w) + bias, 1e-2, "bf16 out")


@pytest.mark.parametrize("T,F,K", SHAPES[:5])
def test_gemm_prefill_f32_resid(T, F, K):
    from paligemma_multimodal_system_b200 import _lib
    x, w = _mk(T, F, K, 1)
    bias = torch.randn(F, device="cuda")
    resid = torch.randn(T, F, device="cuda")
    out = torch.full((T, F), float("nan"), device="cuda")
    _lib.gemm(x, w, out, mode=_lib.EPI_F32, bias=bias, resid=resid, scale=1.0, swap=0)
    torch.cuda.synchronize()
    _close(out, _ref(x, w) + bias + resid, 2e-3, "f32 resid out")
    # in place on the residual stream
    r2 = resid.clone()
    _lib.gemm(x, w, r2, mode=_lib.EPI_F32, bias=bias, resid=r2, swap=0)
    torch.cuda.synchronize()
    _close(r2, _ref(x, w) + bias + resid, 2e-3, "f32 in-place resid")


def test_gemm_gelu():
    from paligemma_multimodal_system_b200 import _lib
    x, w = _mk(384, 4304, 1152, 2)
    bias = torch.randn(4304, device="cuda") * 0.1
    out = torch.empty(384, 4304, device="cuda", dtype=torch.bfloat16)
    _lib.gemm(x, w, out, mode=_lib.EPI_BF16, bias=bias, act_gelu=True, swap=0)
    torch.cuda.synchronize()
    _close(out, torch.nn.functional.gelu(_ref(x, w) + bias, approximate="tanh"), 1e-2, "gelu")


@pytest.mark.parametrize("T", [1, 8, 16, 17, 64, 100, 128])
@pytest.mark.parametrize("F,K", [(2560, 2048), (1281, 256), (2048, 16384)])
def test_gemm_swap_f32(T, F, K):
    from paligemma_multimodal_system_b200 import _lib
    x, w = _mk(T, F, K, 3)
    bias = torch.randn(F, device="cuda")
    out = torch.full((T, F), float("nan"), device="cuda")
    _lib.gemm(x, w, out, mode=_lib.EPI_F32, bias=bias, scale=0.5, swap=1)
    torch.cuda.synchronize()
    _close(out, (_ref(x, w) + bias) * 0.5, 2e-3, "swap f32")
    outb = torch.empty(T, F, device="cuda", dtype=torch.bfloat16)
    _lib.gemm(x, w, outb, mode=_lib.EPI_BF16, swap=1)
    torch.cuda.synchronize()
    _close(outb, _ref(x, w), 1e-2, "swap bf16")


@pytest.mark.parametrize("T,split", [(1, 4), (64, 8), (64, 37), (128, 3)])
def test_gemm_swap_splitk_atomic(T, split):
    from paligemma_multimodal_system_b200 import _lib
    F, K = 2048, 16384
    x, w = _mk(T, F, K, 4)
    resid = torch.randn(T, F, device="cuda")
    out = resid.clone()
    _lib.gemm(x, w, out, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=split)
    torch.cuda.synchronize()
    _close(out, _ref(x, w) + resid, 2e-3, "split-K atomic")


def _pack(gate, up):
    from paligemma_multimodal_system_b200 import _lib
    F, K = gate.shape
    packed = torch.empty(2 * F, K, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().pg_pack_gate_up(gate.data_ptr(), up.data_ptr(), packed.data_ptr(), F, K, _lib.stream()), "pack")
    return packed


@pytest.mark.parametrize("T,swap", [(1, 1), (64, 1), (128, 1), (256, 0), (260, 0), (1000, 0)])
@pytest.mark.parametrize("F,K", [(1024, 256), (16384, 2048)])
def test_gemm_geglu(T, swap, F, K):
    from paligemma_multimodal_system_b200 import _lib
    x, gate = _mk(T, F, K, 5)
    _, up = _mk(T, F, K, 6)
    packed = _pack(gate, up)
    # the packing itself
    ref_packed = torch.stack([gate.view(F // 64, 64, K), up.view(F // 64, 64, K)], 1).reshape(2 * F, K)
    assert torch.equal(packed, ref_packed)
    out = torch.full((T, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.gemm(x, packed, out, mode=_lib.EPI_GEGLU, swap=swap)
    torch.cuda.synchronize()
    ref = torch.nn.functional.gelu(_ref(x, gate), approximate="tanh") * _ref(x, up)
    _close(out, ref, 1.5e-2, "geglu")


def test_gemm_rejects_bad_args():
    from paligemma_multimodal_system_b200 import _lib
    x, w = _mk(16, 64, 60)  # K % 8 != 0
    out = torch.empty(16, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        _lib.gemm(x, w, out, mode=_lib.EPI_BF16)


@pytest.mark.parametrize("T,F,K,resid", [(16500, 1152, 4304, True), (16640, 2048, 6400, True), (16390, 1152, 4304, False)])
def test_gemm_prefill_n_fast_raster_f32(T, F, K, resid):
    """Large activation matrix (> 100 MB) x small weight matrix: the persistent grid walks feature tiles first
    (decode_tile n_fast) and the fp32 epilogue goes through the coalescing transposition tile; token tail not a multiple
    of 128."""
    from paligemma_multimodal_system_b200 import _lib
    x, w = _mk(T, F, K, 11)
    bias = torch.randn(F, device="cuda")
    r = torch.randn(T, F, device="cuda") if resid else None
    out = r.clone() if resid else torch.full((T, F), float("nan"), device="cuda")
    ref = _ref(x, w) + bias + (r if resid else 0)
    _lib.gemm(x, w, out, mode=_lib.EPI_F32, bias=bias, resid=out if resid else None, swap=0)
    torch.cuda.synchronize()
    _close(out, ref, 2e-3, "n-fast f32")


@pytest.mark.parametrize("T,F,K", [(33000, 2560, 2048), (16400, 3456, 1152)])
def test_gemm_prefill_bf16_coalesced_store(T, F, K):
    """bf16 epilogue through the warp transposition tile (full 128-byte lines), with bias + gelu and a token tail."""
    from paligemma_multimodal_system_b200 import _lib
    x, w = _mk(T, F, K, 12)
    bias = torch.randn(F, device="cuda") * 0.2
    out = torch.full((T, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.gemm(x, w, out, mode=_lib.EPI_BF16, bias=bias, act_gelu=True, swap=0)
    torch.cuda.synchronize()
    _close(out, torch.nn.functional.gelu(_ref(x, w) + bias, approximate="tanh"), 1e-2, "bf16 coalesced + gelu")


@pytest.mark.parametrize("B,S,Hq,Hkv,dh,K,cache", [
    (2, 260, 8, 1, 256, 2048, True), (3, 37, 4, 1, 64, 256, True), (1, 4100, 8, 1, 256, 512, True), (2, 130, 4, 2, 64, 256, True),
    (1, 300, 8, 1, 256, 256, False), (5, 64, 2, 2, 256, 128, True)])
def test_qkv_projection_with_rope_and_kv_append_epilogue(B, S, Hq, Hkv, dh, K, cache):
    """pg_gemm_qkv_rope == Linear(q|k|v) -> rotate-half RoPE of the q and k heads (fp32 cos / sin of pos * inv_freq,
    modeling_gemma.py:116-151) -> bf16, k / v rows appended to their cache pages (KVCache.update): one GEMM launch."""
    from paligemma_multimodal_system_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + S)
    T, W = B * S, (Hq + 2 * Hkv) * dh
    x = (torch.randn(T, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(W, K, device="cuda", generator=g) * (1.0 / K ** 0.5)).bfloat16()
    pos = (torch.arange(S, device="cuda").repeat(B) + torch.randint(1, 50, (B,), device="cuda", generator=g).repeat_interleave(S)).int()
    inv_freq = (1.0 / (10000.0 ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))).cuda()
    page, base0 = 64, 70  # the sequences' slots start inside the second page
    max_pages = (base0 + S + page - 1) // page + 1
    pool = B * max_pages + 3
    k_pages = torch.full((pool, page, Hkv * dh), 7.0, device="cuda", dtype=torch.bfloat16)
    v_pages = torch.full((pool, page, Hkv * dh), 7.0, device="cuda", dtype=torch.bfloat16)
    table = torch.randperm(pool, device="cuda", generator=g)[: B * max_pages].int().view(B, max_pages).contiguous()
    slot_base = torch.full((B,), base0, device="cuda", dtype=torch.int32)
    out = torch.full((T, W), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = _lib.lib().pg_gemm_qkv_rope(x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), W, T, K, Hq, Hkv, dh, pos.data_ptr(), inv_freq.data_ptr(),
                                     k_pages.data_ptr() if cache else 0, v_pages.data_ptr() if cache else 0, table.data_ptr() if cache else 0,
                                     slot_base.data_ptr() if cache else 0, S, page, max_pages if cache else 0, _lib.stream())
    _lib.check(rc, "pg_gemm_qkv_rope")
    torch.cuda.synchronize()
    y = _ref(x, w).view(T, Hq + 2 * Hkv, dh)
    ang = pos.float()[:, None] * inv_freq[None, :]                      # fp32 product, as the reference forms it
    cos, sin = torch.cat([ang, ang], -1).cos()[:, None, :], torch.cat([ang, ang], -1).sin()[:, None, :]
    half = dh // 2
    rot = torch.cat([-y[..., half:], y[..., :half]], -1)
    ref = y.clone()
    ref[:, : Hq + Hkv] = y[:, : Hq + Hkv] * cos + rot[:, : Hq + Hkv] * sin
    ref = ref.reshape(T, W)
    _close(out, ref, 1e-2, "qkv + rope epilogue (dense)")
    if cache:
        kv_ref = out.view(T, Hq + 2 * Hkv, dh)
        for b in range(B):
            slots = base0 + torch.arange(S, device="cuda")
            pg_ = table[b, (slots // page).long()].long()
            rows_k = k_pages[pg_, (slots % page).long()].view(S, Hkv, dh)
            rows_v = v_pages[pg_, (slots % page).long()].view(S, Hkv, dh)
            assert torch.equal(rows_k, kv_ref[b * S:(b + 1) * S, Hq: Hq + Hkv]), "k pages hold exactly the dense rotated k rows"
            assert torch.equal(rows_v, kv_ref[b * S:(b + 1) * S, Hq + Hkv:]), "v pages hold exactly the dense v rows"
        # nothing outside the appended slots was touched
        touched = (k_pages != 7.0).any(-1).sum().item()
        assert touched == B * S, (touched, B * S)


# ---------------------------------------------------------------------------------------------------------------------
# CTA-pair kernel (tcgen05.mma.cta_group::2) and banded raster: same arithmetic in the same order as the one-CTA kernel
# ---------------------------------------------------------------------------------------------------------------------
def _with_gemm_mode(mode, fn):
    """mode bit 0: CTA-pair kernel allowed, bit 1: banded raster OFF (pg_debug_set_gemm_pair)."""
    from paligemma_multimodal_system_b200 import _lib
    L = _lib.lib()
    L.pg_debug_set_gemm_pair(mode, 0)
    try:
        return fn()
    finally:
        L.pg_debug_set_gemm_pair(1, 0)


@pytest.mark.parametrize("T,F,K,kind", [
    (19000, 1152, 1152, "f32r"),      # short K: eight epilogue warps, fp32 residual in place, token tail
    (18944 + 77, 3456 + 8, 1152, "bf16b"),  # ragged tokens AND features (TMA zero fill + masked stores)
    (19000, 2048, 4304, "f32r"),      # long K, feature tiles first
    (19200, 8192, 2048, "geglu"),     # gate||up: both operands beyond the L2 budget -> banded raster
    (19000, 2560, 2048, "bf16"),
    (18999, 2048, 1152, "f32"),
])
def test_gemm_cta_pair_bitwise_equal_to_one_cta_kernel(T, F, K, kind):
    """gemm_pair_kernel (256 x 256 tile on a 2-CTA cluster) == gemm_tcgen05_kernel<256> bit for bit, with and without the
    banded tile order, and both within tolerance of the fp32 reference."""
    from paligemma_multimodal_system_b200 import _lib
    x, w = _mk(T, F, K, 21)
    bias = torch.randn(F, device="cuda")
    res0 = torch.randn(T, F, device="cuda") if kind == "f32r" else None

    def run():
        if kind in ("bf16b", "bf16"):
            out = torch.full((T, F), float("nan"), device="cuda", dtype=torch.bfloat16)
            _lib.gemm(x, w, out, mode=_lib.EPI_BF16, bias=bias if kind == "bf16b" else None, swap=0)
        elif kind == "geglu":
            out = torch.full((T, F // 2), float("nan"), device="cuda", dtype=torch.bfloat16)
            _lib.gemm(x, w, out, mode=_lib.EPI_GEGLU, swap=0)
        elif kind == "f32r":
            out = res0.clone()
            _lib.gemm(x, w, out, mode=_lib.EPI_F32, bias=bias, resid=out, swap=0)
        else:
            out = torch.full((T, F), float("nan"), device="cuda")
            _lib.gemm(x, w, out, mode=_lib.EPI_F32, swap=0)
        torch.cuda.synchronize()
        return out

    outs = [_with_gemm_mode(m, run) for m in (2, 3, 0, 1)]  # one-CTA, pair, one-CTA banded, pair banded
    for o in outs[1:]:
        assert torch.equal(outs[0], o)
    acc = _ref(x, w)
    if kind == "bf16b":
        _close(outs[0], acc + bias, 1e-2, "pair bf16+bias")
    elif kind == "bf16":
        _close(outs[0], acc, 1e-2, "pair bf16")
    elif kind == "f32r":
        _close(outs[0], acc + bias + res0, 2e-3, "pair f32 resid")
    elif kind == "f32":
        _close(outs[0], acc, 2e-3, "pair f32")
    else:
        g = acc.view(T, F // 128, 2, 64)
        ref = (torch.nn.functional.gelu(g[:, :, 0], approximate="tanh") * g[:, :, 1]).reshape(T, F // 2)
        _close(outs[0], ref, 2e-2, "pair geglu")


def test_l2_prefetch_never_changes_results_and_validates_arguments():
    """pg_prefetch_l2 only asks the copy engines for cache lines: the GEMM that follows sees the same weights."""
    from paligemma_multimodal_system_b200 import _lib
    L = _lib.lib()
    x, w = _mk(64, 4096, 2048, 5)
    out0 = torch.empty(64, 4096, device="cuda", dtype=torch.bfloat16)
    out1 = torch.empty_like(out0)
    _lib.gemm(x, w, out0, mode=_lib.EPI_BF16, swap=1)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    _lib.check(L.pg_prefetch_l2(w.data_ptr(), w.numel() * 2, 32, 0, side.cuda_stream), "pf")
    _lib.check(L.pg_prefetch_l2(w.data_ptr(), (w.numel() * 2 - 4096) & ~15, 0, 1, side.cuda_stream), "pf")  # ragged length, all SMs, evict-last
    _lib.gemm(x, w, out1, mode=_lib.EPI_BF16, swap=1)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    assert torch.equal(out0, out1)
    assert L.pg_prefetch_l2(0, 1024, 0, 0, _lib.stream()) != 0          # null pointer
    assert L.pg_prefetch_l2(w.data_ptr() + 2, 1024, 0, 0, _lib.stream()) != 0  # misaligned
    assert L.pg_prefetch_l2(w.data_ptr(), 8, 0, 0, _lib.stream()) != 0      # shorter than one 16-byte piece
    assert L.pg_prefetch_l2(w.data_ptr(), 1024, -1, 0, _lib.stream()) != 0  # negative grid
