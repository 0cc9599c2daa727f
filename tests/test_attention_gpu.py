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


@pytest.mark.parametrize("B,Hq,dh,lens", [(4, 8, 256, [261, 300, 64, 1]), (2, 4, 64, [17, 130]), (64, 8, 256, None), (3, 8, 256, [128, 129, 65]),
                                          (1, 8, 256, [4100]), (40, 8, 256, None), (160, 8, 256, None)])
def test_decode_attention_fused_rope_append_combine(B, Hq, dh, lens):
    """pg_attention_decode_fused == RoPE(q,k_new) + cache append + softmax(QK^T/sqrt(dh))V over the whole cache."""
    from paligemma_multimodal_system_b200 import _lib
    if lens is None:
        lens = [260 + (i % 70) for i in range(B)]
    page = 64
    max_pages = (max(lens) + page - 1) // page + 1
    num_pages = B * max_pages + 3
    g = torch.Generator(device="cuda").manual_seed(0)
    k_pages = (torch.randn(num_pages, page, dh, device="cuda", generator=g) * 0.5).bfloat16()
    v_pages = (torch.randn(num_pages, page, dh, device="cuda", generator=g) * 0.5).bfloat16()
    table = torch.randperm(num_pages, device="cuda", generator=g)[: B * max_pages].int().view(B, max_pages).contiguous()
    qkv = torch.randn(B, (Hq + 2) * dh, device="cuda", generator=g) * 0.5
    kv_len = torch.tensor(lens, device="cuda", dtype=torch.int32)
    pos = torch.tensor([l + 3 for l in lens], device="cuda", dtype=torch.int32)
    inv_freq = (1.0 / (10000.0 ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))).cuda()
    out = torch.full((B, Hq * dh), float("nan"), device="cuda", dtype=torch.bfloat16)
    k_before, v_before = k_pages.clone(), v_pages.clone()
    scale = 1.0 / math.sqrt(dh)
    for rep in range(2):
        k_pages.copy_(k_before); v_pages.copy_(v_before)
        rc = _lib.lib().pg_attention_decode_fused(qkv.data_ptr(), pos.data_ptr(), kv_len.data_ptr(), inv_freq.data_ptr(),
                                                  k_pages.data_ptr(), v_pages.data_ptr(), table.data_ptr(), out.data_ptr(),
                                                  B, Hq, 1, dh, page, num_pages, max_pages, scale, _lib.stream())
        _lib.check(rc, "fused decode attn")
        torch.cuda.synchronize()
    half = dh // 2
    rot = lambda t: torch.cat([-t[..., half:], t[..., :half]], -1)
    for b in range(B):
        L = lens[b]
        ang = pos[b].float() * inv_freq
        emb = torch.cat([ang, ang])
        cos, sin = emb.cos(), emb.sin()
        q = qkv[b, : Hq * dh].view(Hq, dh)
        kn = qkv[b, Hq * dh: (Hq + 1) * dh]
        vn = qkv[b, (Hq + 1) * dh:]
        qr = (q * cos + rot(q) * sin).bfloat16().float()
        kr = (kn * cos + rot(kn) * sin).bfloat16()
        idx = table[b].long()
        K = k_before[idx].reshape(-1, dh)[:L].clone()
        V = v_before[idx].reshape(-1, dh)[:L].clone()
        K[L - 1], V[L - 1] = kr, vn.bfloat16()
        ref = torch.softmax(qr @ K.float().t() * scale, -1) @ V.float()
        _close(out[b].view(Hq, dh), ref, 1.5e-2, f"fused decode attention row {b}")
        # the append landed in the right page slot, nothing else in the cache changed
        pg_, off = table[b, (L - 1) // page].item(), (L - 1) % page
        assert (k_pages[pg_, off].float() - kr.float()).abs().max() <= 2 ** -7 * kr.float().abs().max()
        assert torch.equal(v_pages[pg_, off], vn.bfloat16())
    changed = (k_pages != k_before).any(-1).sum().item()
    assert changed <= B


@pytest.mark.parametrize("B,H,N,dh,group", [(1, 16, 4096, 72, 1), (1, 1, 4100, 256, 8), (2, 2, 333, 64, 4), (1, 3, 129, 72, 1)])
def test_prefill_attention_tcgen05_long_and_ragged(B, H, N, dh, group):
    """tcgen05 / TMEM prefill attention at the 896-px lengths (many key tiles: lazy running-max rescale, ring reuse) and
    at lengths that are not multiples of the 128-row / 64-128-key tiles; GQA rows stacked per token (group)."""
    from paligemma_multimodal_system_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(5)
    q = (torch.randn(B, N, H, group, dh, device="cuda", generator=g) * 0.8).bfloat16()
    k = (torch.randn(B, N, H, dh, device="cuda", generator=g) * 0.8).bfloat16()
    v = (torch.randn(B, N, H, dh, device="cuda", generator=g) * 0.8).bfloat16()
    # a few large-magnitude keys late in the sequence force the running maximum to move (O rescale path)
    k[:, N // 2, :, :] *= 6.0
    k[:, N - 3, :, :] *= 9.0
    out = torch.full((B, N, H, group, dh), float("nan"), device="cuda", dtype=torch.bfloat16)
    scale = dh ** -0.5
    rc = _lib.lib().pg_attention_prefill(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, N * group, N, dh, group,
        N * H * group * dh, H * group * dh, dh, group * dh, N * H * dh, H * dh, dh,
        N * H * group * dh, H * group * dh, dh, group * dh, scale, _lib.stream())
    _lib.check(rc, "attn")
    torch.cuda.synchronize()
    qf = q.float().permute(0, 2, 3, 1, 4)                    # [B,H,G,N,dh]
    kf = k.float().permute(0, 2, 1, 3).unsqueeze(2)          # [B,H,1,N,dh]
    vf = v.float().permute(0, 2, 1, 3).unsqueeze(2)
    ref = torch.softmax(qf @ kf.transpose(-1, -2) * scale, -1) @ vf   # [B,H,G,N,dh]
    ref = ref.permute(0, 3, 1, 2, 4)
    _close(out, ref, 2e-2, "tcgen05 prefill attention")


def test_kv_gather_dense_view():
    from paligemma_multimodal_system_b200 import _lib
    B, dh, page, max_pages, n = 3, 256, 64, 4, 150
    g = torch.Generator(device="cuda").manual_seed(1)
    k_pages = (torch.randn(B * max_pages + 2, page, dh, device="cuda", generator=g)).bfloat16()
    perm = torch.randperm(B * max_pages + 2, device="cuda", generator=g)[: B * max_pages].int().view(B, max_pages).contiguous()
    dense = torch.empty(B, 1, n, dh, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().pg_kv_gather(k_pages.data_ptr(), perm.data_ptr(), dense.data_ptr(), B, n, 1, dh, page, max_pages, _lib.stream()), "gather")
    torch.cuda.synchronize()
    for b in range(B):
        assert torch.equal(dense[b, 0], k_pages[perm[b].long()].reshape(-1, dh)[:n])


def test_attention_shapes_without_a_kernel_are_errors():
    """One kernel per entry point: what it cannot express is PG_ERR_ARG, never a silent second implementation."""
    from paligemma_multimodal_system_b200 import _lib
    x = torch.zeros(4096, device="cuda", dtype=torch.bfloat16)
    # dh = 48 has no tcgen05 instantiation; a GQA group of 3 does not tile 128 query rows
    assert _lib.lib().pg_attention_prefill(x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), 1, 1, 8, 8, 48, 1, 384, 48, 0, 48,
                                           384, 48, 48, 384, 48, 0, 48, 0.1, _lib.stream()) == -1
    assert _lib.lib().pg_attention_prefill(x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), 1, 1, 6, 2, 64, 3, 384, 192, 64, 0,
                                           128, 64, 0, 384, 192, 64, 0, 0.1, _lib.stream()) == -1
    f = torch.zeros(4096, device="cuda")
    i = torch.ones(4, device="cuda", dtype=torch.int32)
    # GQA group 16 > 8: no decode kernel
    assert _lib.lib().pg_attention_decode_fused(f.data_ptr(), i.data_ptr(), i.data_ptr(), f.data_ptr(), x.data_ptr(), x.data_ptr(),
                                                i.data_ptr(), x.data_ptr(), 1, 16, 1, 64, 64, 1, 1, 0.1, _lib.stream()) == -1
