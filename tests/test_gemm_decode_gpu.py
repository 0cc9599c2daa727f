"""Cluster split-K decode GEMM (pg_gemm_decode), the RMSNorm-folding epilogues (pg_gemm_bf16_colnorm) and the decode
prologue, each against a plain PyTorch fp32 reference of the same op, through the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(T, F, K, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(T, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(F, K, device="cuda", generator=g) * 0.05).bfloat16()
    return x, w


def _close(a, b, tol, what):
    err = (a.float() - b.float()).abs().max().item()
    ref = b.float().abs().max().item()
    assert err <= tol * max(ref, 1e-6), f"{what}: max-abs err {err:.4g} vs ref absmax {ref:.4g}"


@pytest.mark.parametrize("T", [1, 2, 8, 16, 17, 33, 64, 100, 128])
@pytest.mark.parametrize("F,K,S", [(2560, 2048, 8), (2048, 2048, 16), (2048, 16384, 16), (2048, 16384, 8), (384, 256, 2),
                                   (1281, 256, 4), (2048, 2048, 1), (200, 576, 8)])
def test_gemm_decode_f32(T, F, K, S):
    from paligemma_multimodal_system_b200 import _lib
    x, w = _mk(T, F, K, 1)
    bias = torch.randn(F, device="cuda")
    ss = torch.rand(T, device="cuda") * K + 1.0
    out = torch.full((T, F), float("nan"), device="cuda")
    _lib.gemm_decode(x, w, out, mode=_lib.DEC_F32, cluster_k=S, bias=bias, ss_in=ss, norm_dim=K, eps=1e-6)
    torch.cuda.synchronize()
    ref = (x.float() @ w.float().t()) * torch.rsqrt(ss / K + 1e-6)[:, None] + bias
    _close(out, ref, 2e-3, "decode f32")
    out2 = torch.full((T, F), float("nan"), device="cuda")
    _lib.gemm_decode(x, w, out2, mode=_lib.DEC_F32, cluster_k=S)
    torch.cuda.synchronize()
    _close(out2, x.float() @ w.float().t(), 2e-3, "decode f32 plain")


@pytest.mark.parametrize("T", [1, 5, 16, 64, 128])
@pytest.mark.parametrize("F,K,S", [(2048, 2048, 8), (2048, 16384, 16), (2048, 16384, 8), (256, 1024, 4), (256, 256, 2)])
def test_gemm_decode_resid_norm(T, F, K, S):
    from paligemma_multimodal_system_b200 import _lib
    x, w = _mk(T, F, K, 2)
    h0 = torch.randn(T, F, device="cuda") * 3
    nw = torch.randn(F, device="cuda") * 0.1
    h = h0.clone()
    hb = torch.full((T, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    ss = torch.zeros(T, device="cuda")
    _lib.gemm_decode(x, w, h, mode=_lib.DEC_RESID_NORM, cluster_k=S, hb=hb, norm_w=nw, ss_out=ss)
    torch.cuda.synchronize()
    ref = h0 + x.float() @ w.float().t()
    _close(h, ref, 2e-3, "residual stream")
    _close(hb, ref * (1 + nw), 1e-2, "bf16 norm operand")
    _close(ss, (ref * ref).sum(-1), 2e-3, "sum of squares")


@pytest.mark.parametrize("T", [1, 8, 16, 31, 64, 128])
@pytest.mark.parametrize("F,K", [(1024, 256), (16384, 2048)])
def test_gemm_geglu_colnorm(T, F, K):
    from paligemma_multimodal_system_b200 import _lib
    x, gate = _mk(T, F, K, 5)
    _, up = _mk(T, F, K, 6)
    packed = torch.empty(2 * F, K, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().pg_pack_gate_up(gate.data_ptr(), up.data_ptr(), packed.data_ptr(), F, K, _lib.stream()), "pack")
    ss = torch.rand(T, device="cuda") * K * 4 + 1.0
    out = torch.full((T, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.gemm_colnorm(x, packed, out, mode=_lib.EPI_GEGLU, ss_in=ss, norm_dim=K)
    torch.cuda.synchronize()
    rs = torch.rsqrt(ss / K + 1e-6)[:, None]
    ref = torch.nn.functional.gelu((x.float() @ gate.float().t()) * rs, approximate="tanh") * ((x.float() @ up.float().t()) * rs)
    _close(out, ref, 1.5e-2, "geglu colnorm")


@pytest.mark.parametrize("T", [1, 64])
def test_gemm_head_colnorm(T):
    from paligemma_multimodal_system_b200 import _lib
    F, K = 1281, 256
    x, w = _mk(T, F, K, 7)
    bias = torch.randn(F, device="cuda")
    ss = torch.rand(T, device="cuda") * K + 1.0
    out = torch.full((T, F), float("nan"), device="cuda")
    _lib.gemm_colnorm(x, w, out, mode=_lib.EPI_F32, bias=bias, ss_in=ss, norm_dim=K)
    torch.cuda.synchronize()
    ref = (x.float() @ w.float().t()) * torch.rsqrt(ss / K + 1e-6)[:, None] + bias
    _close(out, ref, 2e-3, "head colnorm")


def test_decode_prologue():
    from paligemma_multimodal_system_b200 import _lib
    B, D, V, N = 5, 256, 300, 4
    g = torch.Generator(device="cuda").manual_seed(3)
    embed = torch.randn(V, D, device="cuda", generator=g).bfloat16()
    img = torch.randn(B, N, D, device="cuda", generator=g)
    nw = torch.randn(D, device="cuda", generator=g) * 0.1
    tok = torch.tensor([7, 0, 299, 256, 12], device="cuda", dtype=torch.int32)  # 0 = pad, 256 = image token
    h = torch.full((B, D), float("nan"), device="cuda")
    hb = torch.empty(B, D, device="cuda", dtype=torch.bfloat16)
    ss = torch.full((4, B), 7.0, device="cuda")
    _lib.check(_lib.lib().pg_decode_prologue(tok.data_ptr(), embed.data_ptr(), img.data_ptr(), h.data_ptr(), hb.data_ptr(),
                                             ss[0].data_ptr(), nw.data_ptr(), ss[1].data_ptr(), 3 * B, B, D, N, 16.0, 0.5, 0, 256,
                                             _lib.stream()), "prologue")
    torch.cuda.synchronize()
    ref = embed[tok.long()].float() * 16.0
    ref[1] = 0
    ref[3] = img[3, 0] * 0.5
    assert torch.allclose(h, ref, rtol=1e-6, atol=1e-6)
    _close(hb, ref * (1 + nw), 1e-2, "hb")
    _close(ss[0], (ref * ref).sum(-1), 1e-5, "ss")
    assert torch.equal(ss[1:], torch.zeros(3, B, device="cuda"))
    # tokens == NULL keeps h
    h2 = torch.randn(B, D, device="cuda")
    keep = h2.clone()
    _lib.check(_lib.lib().pg_decode_prologue(0, 0, 0, h2.data_ptr(), hb.data_ptr(), ss[0].data_ptr(), nw.data_ptr(), 0, 0, B, D, 0,
                                             1.0, 1.0, -1, -1, _lib.stream()), "prologue")
    torch.cuda.synchronize()
    assert torch.equal(h2, keep)
    _close(hb, keep * (1 + nw), 1e-2, "hb from h")
    _close(ss[0], (keep * keep).sum(-1), 1e-5, "ss from h")
