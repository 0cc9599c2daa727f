"""Decode-step chores of the swap-AB GEMM (pg_gemm_bf16_fused): the zero-fill of a later split-K accumulator that rides on the
o_proj launch -- against plain PyTorch fp32 of the same math, through the C ABI.  (The lm_head statistics epilogue is covered
by tests/test_sampler_gpu.py.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    err = (a.float() - b.float()).abs().max().item()
    ref = b.float().abs().max().item()
    assert err <= tol * max(ref, 1e-6), f"{what}: max-abs err {err:.4g} vs ref absmax {ref:.4g}"


@pytest.mark.parametrize("T,split", [(1, 9), (16, 4), (40, 9), (64, 9), (128, 2)])
def test_zero_fill_rides_along(T, split):
    """o_proj of a decode step: split-K red.add into the residual stream, re-zeroing the q/k/v accumulator (every token-tile
    width of the swap kernels: 16 / 32 / 64 with eight epilogue warps / 128)."""
    from paligemma_multimodal_system_b200 import _lib
    F, K = 2048, 2048
    g = torch.Generator(device="cuda").manual_seed(4)
    x = (torch.randn(T, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(F, K, device="cuda", generator=g) * 0.05).bfloat16()
    resid = torch.randn(T, F, device="cuda", generator=g)
    out = resid.clone()
    acc = torch.full((T, 2560), 3.0, device="cuda")
    _lib.gemm_fused(x, w, out, mode=_lib.EPI_ATOMIC_F32, split_k=split, zero_buf=acc)
    torch.cuda.synchronize()
    _close(out, x.float() @ w.float().t() + resid, 2e-3, "o_proj split-K")
    assert torch.count_nonzero(acc) == 0


def test_fused_rejects_bad_args():
    from paligemma_multimodal_system_b200 import _lib
    x = torch.zeros(8, 256, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(256, 256, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(8, 256, device="cuda")
    with pytest.raises(RuntimeError):  # zero-fill granularity (float4)
        _lib.gemm_fused(x, w, out, mode=_lib.EPI_ATOMIC_F32, zero_buf=torch.zeros(6, device="cuda"))
    with pytest.raises(RuntimeError):  # statistics belong to the plain fp32 logits epilogue
        _lib.gemm_fused(x, w, out, mode=_lib.EPI_ATOMIC_F32, stats=torch.zeros(8, 8, 2, device="cuda"))
