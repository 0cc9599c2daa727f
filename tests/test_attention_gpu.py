"""Attention kernels against plain PyTorch fp32 softmax(QK^T*scale)V, through the C ABI."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    err = (a.float() - b.float()).abs().max().item()
    ref = b.float().abs().max().item()
    assert err <= tol * max(ref, 1e-6), f"{what}: max-abs err {err:.4g} vs ref absmax {ref:.4g}"


@pytest.mark.parametrize("B,H,N,dh", [(2, 16, 256, 72), (1, 4, 256, 64), (2, 16, 1024, 72), (1, 3, 100, 72), (1, 2, 65, 64)])
def test_siglip_attention(B, H, N, dh):
    """q/k/v are column slices of one fused [B*N, 3*H*dh] projection buffer (modeling_siglip.py:71-136)."""
    from paligemma_multimodal_system_b200 import _lib
    D = H * dh
    qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.7).bfloat16()
    out = torch.full((B * N, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    scale = dh ** -0.5
    rc = _lib.lib().pg_attention_prefill(
        qkv.data_ptr(), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, out.data_ptr(), B, H, N, N, dh, 1,
        N * 3 * D, 3 * D, 0, dh, N * 3 * D, 3 * D, dh, N * D, D, 0, dh, scale, _lib.stream())
    _lib.check(rc, "attn")
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    ref = torch.softmax(q @ k.transpose(-1, -2) * scale, -1) @ v
    ref = ref.permute(0, 2, 1, 3).reshape(B * N, D)
    _close(out, ref, 1.5e-2, "siglip attention")


@pytest.mark.parametrize("B,S,Hq,dh", [(2, 260, 8, 256), (1, 1028, 8, 256), (3, 37, 4, 64), (1, 300, 4, 64)])
def test_gemma_prefill_attention(B, S, Hq, dh):
    """MQA: the Hq heads of one token are consecutive query rows against a single KV head (modeling_gemma.py:285-339)."""
    from paligemma_multimodal_system_b200 import _lib
    q = (torch.randn(B * S, Hq * dh, device="cuda") * 0.5).bfloat16()
    k = (torch.randn(B * S, dh, device="cuda") * 0.5).bfloat16()
    v = (torch.randn(B * S, dh, device="cuda") * 0.5).bfloat16()
    out = torch.full((B * S, Hq * dh), float("nan"), device="cuda", dtype=torch.bfloat16)
    scale = 1.0 / math.sqrt(dh)
    rc = _lib.lib().pg_attention_prefill(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, 1, S * Hq, S, dh, Hq,
        S * Hq * dh, Hq * dh, dh, 0, S * dh, dh, 0, S * Hq * dh, Hq * dh, dh, 0, scale, _lib.stream())
    _lib.check(rc, "attn")
    torch.cuda.synchronize()
    qf = q.float().view(B, S, Hq, dh).transpose(1, 2)
    kf = k.float().view(B, 1, S, dh)
    vf = v.float().view(B, 1, S, dh)
    ref = torch.softmax(qf @ kf.transpose(-1, -2) * scale, -1) @ vf
    ref = ref.transpose(1, 2).reshape(B * S, Hq * dh)
    _close(out, ref, 1.5e-2, "gemma prefill attention")


@pytest.mark.parametrize("B,Hq,dh,lens,splits", [
    (4, 8, 256, [261, 300, 64, 1], 4), (2, 4, 64, [17, 130], 2), (64, 8, 256, None, 4), (1, 8, 256, [700], 8),
    (3, 8, 256, [128, 129, 127], 1)])
def test_decode_attention_paged(B, Hq, dh, lens, splits):
    from paligemma_multimodal_system_b200 import _lib
    if lens is None:
        lens = [260 + (i % 7) for i in range(B)]
    page = 64
    max_pages = (max(lens) + page - 1) // page + 1
    num_pages = B * max_pages + 3
    g = torch.Generator(device="cuda").manual_seed(0)
    k_pages = (torch.randn(num_pages, page, dh, device="cuda", generator=g) * 0.5).bfloat16()
    v_pages = (torch.randn(num_pages, page, dh, device="cuda", generator=g) * 0.5).bfloat16()
    perm = torch.randperm(num_pages, device="cuda", generator=g)[: B * max_pages].int().view(B, max_pages).contiguous()
    q = (torch.randn(B, Hq * dh, device="cuda", generator=g) * 0.5).bfloat16()
    kv_len = torch.tensor(lens, device="cuda", dtype=torch.int32)
    out = torch.full((B, Hq * dh), float("nan"), device="cuda", dtype=torch.bfloat16)
    ws = torch.empty(_lib.lib().pg_attention_decode_workspace_floats(B, Hq, dh, splits), device="cuda")
    scale = 1.0 / math.sqrt(dh)
    rc = _lib.lib().pg_attention_decode(q.data_ptr(), k_pages.data_ptr(), v_pages.data_ptr(), perm.data_ptr(), kv_len.data_ptr(),
                                        out.data_ptr(), ws.data_ptr(), B, Hq, 1, dh, page, max_pages, splits, scale, _lib.stream())
    _lib.check(rc, "decode attn")
    torch.cuda.synchronize()
    for b in range(B):
        L = lens[b]
        idx = perm[b].long()
        K = k_pages[idx].reshape(-1, dh)[:L].float()
        V = v_pages[idx].reshape(-1, dh)[:L].float()
        qb = q[b].float().view(Hq, dh)
        ref = torch.softmax(qb @ K.t() * scale, -1) @ V
        _close(out[b].view(Hq, dh), ref, 1.5e-2, f"decode attention row {b}")

    # gather view
    dense = torch.empty(B, 1, min(lens), dh, device="cuda", dtype=torch.bfloat16)
    rc = _lib.lib().pg_kv_gather(k_pages.data_ptr(), perm.data_ptr(), dense.data_ptr(), B, min(lens), 1, dh, page, max_pages, _lib.stream())
    _lib.check(rc, "gather")
    torch.cuda.synchronize()
    for b in range(B):
        assert torch.equal(dense[b, 0], k_pages[perm[b].long()].reshape(-1, dh)[: min(lens)])
