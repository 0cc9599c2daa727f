"""Decode-step GEMM fusions (pg_gemm_bf16_fused): the activation operand built in the kernel from the fp32 residual stream
(GemmaRMSNorm folded into the projection, modeling_gemma.py:172-181,395-396,412-413), the zero-fill of a later split-K
accumulator -- against plain PyTorch fp32 of the same math, through the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    err = (a.float() - b.float()).abs().max().item()
    ref = b.float().abs().max().item()
    assert err <= tol * max(ref, 1e-6), f"{what}: max-abs err {err:.4g} vs ref absmax {ref:.4g}"


def _mk(T, F, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    h = torch.randn(T, K, device="cuda", generator=g) * torch.linspace(0.2, 4.0, T, device="cuda")[:, None]  # rows of very different norms
    nw = torch.randn(K, device="cuda", generator=g) * 0.2
    w = (torch.randn(F, K, device="cuda", generator=g) * 0.05).bfloat16()
    return h, nw, w


def _operand(h, nw):
    """What the kernel feeds the tensor core: bf16(h * (1 + w)), one rounding."""
    return (h * (1.0 + nw)).bfloat16().float()


@pytest.mark.parametrize("T", [1, 5, 16, 17, 32, 33, 64, 100, 128])
@pytest.mark.parametrize("F,K,split", [(2560, 2048, 7), (1536, 256, 2), (640, 200, 1)])
def test_fused_operand_splitk_atomic(T, F, K, split):
    """q/k/v projection of a decode step: split-K red.add of Linear(bf16(h * (1 + w))) into a zeroed accumulator; the
    per-token RMSNorm factor is left to the consumer.  Every token-tile width (16/32/64/64x8 warps/128) and a K tail."""
    from paligemma_multimodal_system_b200 import _lib
    h, nw, w = _mk(T, F, K, 1)
    out = torch.zeros(T, F, device="cuda")
    _lib.gemm_fused(w, out, mode=_lib.EPI_ATOMIC_F32, x_f32=h, norm_w=nw, split_k=split)
    torch.cuda.synchronize()
    _close(out, _operand(h, nw) @ w.float().t(), 2e-3, "fused operand, split-K")
    # a strided view of a wider residual buffer (row pitch != K)
    wide = torch.zeros(T, K + 64, device="cuda")
    wide[:, :K] = h
    out2 = torch.zeros(T, F, device="cuda")
    _lib.gemm_fused(w, out2, mode=_lib.EPI_ATOMIC_F32, x_f32=wide[:, :K], norm_w=nw, split_k=split)
    torch.cuda.synchronize()
    _close(out2, _operand(h, nw) @ w.float().t(), 2e-3, "fused operand, row pitch")


@pytest.mark.parametrize("T", [1, 7, 16, 32, 40, 64, 128])
@pytest.mark.parametrize("F,K", [(1024, 256), (16384, 2048)])
def test_fused_rmsnorm_geglu(T, F, K):
    """gate||up of a decode step: out = gelu_tanh(gate(n)) * up(n) with n = GemmaRMSNorm(h); operand from the fp32 rows, the
    factor rsqrt(mean(h^2) + eps) computed in the kernel and applied to the accumulator."""
    from paligemma_multimodal_system_b200 import _lib
    from test_gemm_gpu import _pack
    h, nw, gate = _mk(T, F, K, 2)
    up = (torch.randn(F, K, device="cuda") * 0.05).bfloat16()
    packed = _pack(gate, up)
    out = torch.full((T, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.gemm_fused(packed, out, mode=_lib.EPI_GEGLU, x_f32=h, norm_w=nw, apply_rstd=True, eps=1e-6)
    torch.cuda.synchronize()
    r = torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + 1e-6)
    x = _operand(h, nw)
    ref = torch.nn.functional.gelu((x @ gate.float().t()) * r, approximate="tanh") * ((x @ up.float().t()) * r)
    _close(out, ref, 1.5e-2, "fused rmsnorm + geglu")
    # and against the literal module chain (fp32 norm, bf16 rounding of its output as the unfused path did)
    n = (h * r * (1.0 + nw)).bfloat16().float()
    ref2 = torch.nn.functional.gelu(n @ gate.float().t(), approximate="tanh") * (n @ up.float().t())
    _close(out, ref2, 2.5e-2, "fused rmsnorm + geglu vs module chain")


@pytest.mark.parametrize("T", [3, 64, 128])
def test_fused_rmsnorm_f32_and_bf16_epilogues(T):
    from paligemma_multimodal_system_b200 import _lib
    F, K = 1281, 256
    h, nw, w = _mk(T, F, K, 3)
    bias = torch.randn(F, device="cuda")
    r = torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + 1e-6)
    ref = (_operand(h, nw) @ w.float().t()) * r + bias
    out = torch.full((T, F), float("nan"), device="cuda")
    _lib.gemm_fused(w, out, mode=_lib.EPI_F32, x_f32=h, norm_w=nw, apply_rstd=True, bias=bias)
    torch.cuda.synchronize()
    _close(out, ref, 2e-3, "fused rmsnorm, fp32 epilogue")
    outb = torch.full((T, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.gemm_fused(w, outb, mode=_lib.EPI_BF16, x_f32=h, norm_w=nw, apply_rstd=True, bias=bias)
    torch.cuda.synchronize()
    _close(outb, ref, 1e-2, "fused rmsnorm, bf16 epilogue")


def test_zero_fill_rides_along():
    """o_proj of a decode step: split-K red.add into the residual stream from a bf16 operand, re-zeroing the q/k/v accumulator."""
    from paligemma_multimodal_system_b200 import _lib
    T, F, K = 64, 2048, 2048
    g = torch.Generator(device="cuda").manual_seed(4)
    x = (torch.randn(T, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(F, K, device="cuda", generator=g) * 0.05).bfloat16()
    resid = torch.randn(T, F, device="cuda", generator=g)
    out = resid.clone()
    acc = torch.full((T, 2560), 3.0, device="cuda")
    _lib.gemm_fused(w, out, mode=_lib.EPI_ATOMIC_F32, x=x, split_k=9, zero_buf=acc)
    torch.cuda.synchronize()
    _close(out, x.float() @ w.float().t() + resid, 2e-3, "o_proj split-K")
    assert torch.count_nonzero(acc) == 0


def test_fused_rejects_bad_args():
    from paligemma_multimodal_system_b200 import _lib
    h, nw, w = _mk(8, 256, 256, 6)
    out = torch.zeros(8, 256, device="cuda")
    with pytest.raises(RuntimeError):  # a split only sees a slice of the row: no in-kernel factor
        _lib.gemm_fused(w, out, mode=_lib.EPI_ATOMIC_F32, x_f32=h, norm_w=nw, apply_rstd=True, split_k=2)
    with pytest.raises(RuntimeError):  # zero-fill granularity
        _lib.gemm_fused(w, out, mode=_lib.EPI_ATOMIC_F32, x_f32=h, norm_w=nw, zero_buf=torch.zeros(6, device="cuda"))
