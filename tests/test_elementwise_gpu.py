"""Row kernels (norms, im2col, merge, RoPE + KV append) against plain PyTorch fp32, through the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    err = (a.float() - b.float()).abs().max().item()
    ref = b.float().abs().max().item()
    assert err <= tol * max(ref, 1e-6), f"{what}: max-abs err {err:.4g} vs ref absmax {ref:.4g}"


@pytest.mark.parametrize("rows,D", [(5, 256), (300, 1152), (64, 2048)])
def test_layernorm(rows, D):
    from paligemma_multimodal_system_b200 import _lib
    x = torch.randn(rows, D, device="cuda") * 3 + 1
    g, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    yb = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    yf = torch.empty(rows, D, device="cuda")
    _lib.layernorm(x, g, b, 1e-6, yb, yf)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-6)
    _close(yf, ref, 1e-5, "layernorm f32")
    _close(yb, ref, 5e-3, "layernorm bf16")


@pytest.mark.parametrize("rows,D", [(1, 256), (64, 2048), (333, 2048)])
def test_rmsnorm(rows, D):
    from paligemma_multimodal_system_b200 import _lib
    x = torch.randn(rows, D, device="cuda") * 5
    w = torch.randn(D, device="cuda") * 0.1
    y = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    _lib.rmsnorm(x, w, y, 1e-6)
    torch.cuda.synchronize()
    ref = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * (1.0 + w)
    _close(y, ref, 5e-3, "rmsnorm")
    assert (y.float() - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-6  # one bf16 rounding


def test_im2col_matches_conv2d():
    from paligemma_multimodal_system_b200 import _lib
    B, C, H, P, Dv = 2, 3, 224, 14, 64
    px = torch.randn(B, C, H, H, device="cuda").bfloat16().float()
    Kp = 592
    patches = torch.empty(B * 256, Kp, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().pg_im2col(px.data_ptr(), patches.data_ptr(), B, C, H, H, P, Kp, _lib.stream()), "im2col")
    torch.cuda.synchronize()
    w = torch.randn(Dv, C, P, P, device="cuda").bfloat16().float()
    ref = torch.nn.functional.conv2d(px, w, stride=P).flatten(2).transpose(1, 2).reshape(B * 256, Dv)
    got = patches[:, : C * P * P].float() @ w.view(Dv, -1).t()
    _close(got, ref, 1e-4, "im2col")
    assert torch.count_nonzero(patches[:, C * P * P:]) == 0


@pytest.mark.parametrize("B,N,D,K", [(3, 256, 1152, 640), (2, 1024, 256, 592), (1, 300, 64, 64)])
def test_patch_gemm_adds_position_embeddings_in_its_epilogue(B, N, D, K):
    """x[b*N + n, :] = patches @ W^T + bias + pos[n, :] (modeling_siglip.py:258-263,289-298): the position-embedding add is the
    fp32 `residual` of the patch GEMM, its row taken modulo the number of patches."""
    from paligemma_multimodal_system_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(2)
    patches = (torch.randn(B * N, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(D, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(D, device="cuda", generator=g)
    pos = torch.randn(N, D, device="cuda", generator=g)
    out = torch.full((B * N, D), float("nan"), device="cuda")
    _lib.gemm(patches, w, out, mode=_lib.EPI_F32, bias=bias, resid=pos, resid_row_mod=N, swap=0)
    torch.cuda.synchronize()
    ref = ((patches.float() @ w.float().t() + bias).view(B, N, D) + pos).view(B * N, D)
    _close(out, ref, 2e-3, "patch GEMM + bias + position embeddings")


def test_merge_embeddings_and_positions():
    from paligemma_multimodal_system_b200 import _lib
    B, N, T, D, V = 3, 256, 6, 256, 1281
    S = N + T
    img_tok, pad = 1024, 0
    g = torch.Generator(device="cuda").manual_seed(1)
    ids = torch.randint(1, 1024, (B, S), device="cuda", generator=g)
    ids[:, :N] = img_tok
    ids[1, -2:] = pad
    # row 2: image tokens not contiguous (still exactly N of them)
    ids[2, 5], ids[2, N] = ids[2, N].item(), img_tok
    mask = (ids != pad).long()
    embed = torch.randn(V, D, device="cuda").bfloat16()
    img = torch.randn(B, N, D, device="cuda")
    h = torch.full((B, S, D), float("nan"), device="cuda")
    pos = torch.empty(B, S, device="cuda", dtype=torch.int32)
    src = torch.empty(B, S, device="cuda", dtype=torch.int32)
    err = torch.zeros(1, device="cuda", dtype=torch.int32)
    ts, isc = D ** 0.5, (D ** -0.5) * (D ** 0.5)
    rc = _lib.lib().pg_merge_embeddings(ids.data_ptr(), mask.data_ptr(), embed.data_ptr(), img.data_ptr(), h.data_ptr(),
                                        pos.data_ptr(), src.data_ptr(), err.data_ptr(), B, S, D, N, img_tok, pad, ts, isc, _lib.stream())
    _lib.check(rc, "merge")
    torch.cuda.synchronize()
    assert err.item() == 0
    # reference semantics (modeling_paligemma.py:99-128 + modeling_gemma.py:510-511)
    e = embed.float()[ids]
    final = torch.zeros(B, S, D, device="cuda")
    tm = ((ids != img_tok) & (ids != pad))[..., None].expand(-1, -1, D)
    im = (ids == img_tok)[..., None].expand(-1, -1, D)
    pm = (ids == pad)[..., None].expand(-1, -1, D)
    final = torch.where(tm, e, final)
    final = final.masked_scatter(im, img * (D ** -0.5))
    final = torch.where(pm, torch.zeros_like(final), final) * (D ** 0.5)
    _close(h, final, 1e-6, "merge")
    refpos = mask.cumsum(-1).masked_fill(mask == 0, 1)
    assert torch.equal(pos.long(), refpos)
    # a row with the wrong number of image tokens raises the flag
    ids[0, 0] = 5
    rc = _lib.lib().pg_merge_embeddings(ids.data_ptr(), mask.data_ptr(), embed.data_ptr(), img.data_ptr(), h.data_ptr(),
                                        pos.data_ptr(), src.data_ptr(), err.data_ptr(), B, S, D, N, img_tok, pad, ts, isc, _lib.stream())
    torch.cuda.synchronize()
    assert err.item() == 1


@pytest.mark.parametrize("f32in", [False, True])
def test_rope_kv_append(f32in):
    from paligemma_multimodal_system_b200 import _lib
    B, Sq, Hq, Hkv, dh, page = 2, 70, 8, 1, 256, 64
    T = B * Sq
    W = (Hq + 2 * Hkv) * dh
    qkv = torch.randn(T, W, device="cuda")
    qkv = qkv if f32in else qkv.bfloat16()
    pos = (torch.arange(Sq, device="cuda") + 1).repeat(B).int().contiguous()
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))
    inv_freq = inv_freq.cuda()
    max_pages = 3
    table = torch.tensor([[4, 1, 5], [0, 3, 2]], device="cuda", dtype=torch.int32)
    slot_base = torch.tensor([10, 0], device="cuda", dtype=torch.int32)
    kp = torch.zeros(6, page, Hkv * dh, device="cuda", dtype=torch.bfloat16)
    vp = torch.zeros_like(kp)
    qo = torch.empty(T, Hq * dh, device="cuda", dtype=torch.bfloat16)
    ko = torch.empty(T, Hkv * dh, device="cuda", dtype=torch.bfloat16)
    vo = torch.empty_like(ko)
    rc = _lib.lib().pg_rope_kv_append(qkv.data_ptr(), int(f32in), pos.data_ptr(), qo.data_ptr(), ko.data_ptr(), vo.data_ptr(),
                                      kp.data_ptr(), vp.data_ptr(), table.data_ptr(), slot_base.data_ptr(), B, Sq, Hq, Hkv, dh,
                                      page, max_pages, inv_freq.data_ptr(), _lib.stream())
    _lib.check(rc, "rope")
    torch.cuda.synchronize()
    x = qkv.float()
    q, k, v = x[:, : Hq * dh].view(T, Hq, dh), x[:, Hq * dh: (Hq + Hkv) * dh].view(T, Hkv, dh), x[:, (Hq + Hkv) * dh:]
    ang = pos.float()[:, None] * inv_freq[None, :]
    emb = torch.cat([ang, ang], -1)
    cos, sin = emb.cos()[:, None, :], emb.sin()[:, None, :]
    rot = lambda t: torch.cat([-t[..., dh // 2:], t[..., : dh // 2]], -1)
    qr, kr = q * cos + rot(q) * sin, k * cos + rot(k) * sin
    _close(qo.view(T, Hq, dh), qr, 1e-2, "q rope")
    _close(ko.view(T, Hkv, dh), kr, 1e-2, "k rope")
    assert torch.equal(vo, v.bfloat16())
    for b in range(B):
        for s in range(Sq):
            slot = slot_base[b].item() + s
            pg_, off = table[b, slot // page].item(), slot % page
            assert torch.equal(kp[pg_, off], ko[b * Sq + s])
            assert torch.equal(vp[pg_, off], vo[b * Sq + s])


@pytest.mark.parametrize("rows,D", [(1025, 256), (16384, 1152), (4099, 2048)])
def test_layernorm_warp_per_row(rows, D):
    """rows >= 1024 and D % 128 == 0 dispatch to the register-resident warp-per-row kernel (prefill)."""
    from paligemma_multimodal_system_b200 import _lib
    x = torch.randn(rows, D, device="cuda") * 3 + 1
    g, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    yb = torch.full((rows, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    yf = torch.full((rows, D), float("nan"), device="cuda")
    _lib.layernorm(x, g, b, 1e-6, yb, yf)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-6)
    _close(yf, ref, 1e-5, "layernorm f32 (warp per row)")
    _close(yb, ref, 5e-3, "layernorm bf16 (warp per row)")


@pytest.mark.parametrize("rows,D", [(1024, 256), (16640, 2048), (1031, 1152)])
def test_rmsnorm_warp_per_row(rows, D):
    from paligemma_multimodal_system_b200 import _lib
    x = torch.randn(rows, D, device="cuda") * 5
    w = torch.randn(D, device="cuda") * 0.1
    y = torch.full((rows, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.rmsnorm(x, w, y, 1e-6)
    torch.cuda.synchronize()
    ref = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * (1.0 + w)
    assert (y.float() - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-6  # one bf16 rounding


def test_projector_rows_scattered_by_the_gemm_epilogue():
    """pg_merge_scan + projector GEMM with out_row_map + pg_merge_text == pg_merge_embeddings on materialised image features
    (_merge_input_ids_with_image_features, modeling_paligemma.py:201-251): ragged image-token positions, pads, text."""
    from paligemma_multimodal_system_b200 import _lib
    L = _lib.lib()
    B, N, T, D, Dv, V = 3, 256, 7, 256, 128, 1281
    S = N + T
    img_tok, pad = 1024, 0
    g = torch.Generator(device="cuda").manual_seed(4)
    ids = torch.randint(1, 1024, (B, S), device="cuda", generator=g)
    ids[:, :N] = img_tok
    ids[1] = torch.cat([ids[1, N:N + 3], torch.full((N,), img_tok, device="cuda"), ids[1, N + 3:]])  # image tokens not at the front
    ids[2, -2:] = pad
    mask = (ids != pad).long()
    embed = (torch.randn(V, D, device="cuda", generator=g)).bfloat16()
    feats = (torch.randn(B * N, Dv, device="cuda", generator=g) * 0.5).bfloat16()
    wp = (torch.randn(D, Dv, device="cuda", generator=g) * 0.1).bfloat16()
    bias = torch.randn(D, device="cuda", generator=g)
    text_scale, img_scale = D ** 0.5, 0.93
    # reference: materialise the projected features, then the one-call merge
    img = torch.empty(B * N, D, device="cuda")
    _lib.gemm(feats, wp, img, mode=_lib.EPI_F32, bias=bias, swap=0)
    h_ref = torch.empty(B * S, D, device="cuda")
    pos_ref = torch.empty(B * S, device="cuda", dtype=torch.int32)
    src = torch.empty(B * S, device="cuda", dtype=torch.int32)
    err = torch.zeros(1, device="cuda", dtype=torch.int32)
    _lib.check(L.pg_merge_embeddings(ids.data_ptr(), mask.data_ptr(), embed.data_ptr(), img.data_ptr(), h_ref.data_ptr(), pos_ref.data_ptr(),
                                     src.data_ptr(), err.data_ptr(), B, S, D, N, img_tok, pad, text_scale, img_scale, _lib.stream()), "merge")
    # fused: scan -> projector scatters -> text rows
    h = torch.full((B * S + 1, D), float("nan"), device="cuda")
    pos = torch.empty(B * S, device="cuda", dtype=torch.int32)
    src2 = torch.empty(B * S, device="cuda", dtype=torch.int32)
    dst = torch.full((B * N,), B * S, device="cuda", dtype=torch.int32)
    err2 = torch.zeros(1, device="cuda", dtype=torch.int32)
    _lib.check(L.pg_merge_scan(ids.data_ptr(), mask.data_ptr(), pos.data_ptr(), src2.data_ptr(), dst.data_ptr(), err2.data_ptr(), B, S, N,
                               img_tok, pad, _lib.stream()), "scan")
    _lib.gemm(feats, wp, h, mode=_lib.EPI_F32, bias=bias, scale=img_scale, swap=0, out_row_map=dst)
    _lib.check(L.pg_merge_text(ids.data_ptr(), src2.data_ptr(), embed.data_ptr(), h.data_ptr(), B, S, D, N, text_scale, _lib.stream()), "text")
    torch.cuda.synchronize()
    assert int(err.item()) == 0 and int(err2.item()) == 0
    assert torch.equal(pos, pos_ref) and torch.equal(src, src2)
    assert sorted(dst.tolist()) == sorted((ids.view(-1) == img_tok).nonzero().view(-1).tolist())
    assert not torch.isnan(h[: B * S]).any()
    assert (h[: B * S] - h_ref).abs().max().item() <= 1e-5 * h_ref.abs().max().item()  # (acc + bias) * s vs ((acc + bias)) * s: one rounding
    # a row with too few image tokens raises the flag; its orphan features go to the sink row
    ids[0, 5] = 7
    dst.fill_(B * S)
    _lib.check(L.pg_merge_scan(ids.data_ptr(), mask.data_ptr(), pos.data_ptr(), src2.data_ptr(), dst.data_ptr(), err2.data_ptr(), B, S, N,
                               img_tok, pad, _lib.stream()), "scan")
    torch.cuda.synchronize()
    assert int(err2.item()) == 1 and int((dst == B * S).sum()) == 1
