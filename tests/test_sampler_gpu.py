"""Sampling kernels against torch.argmax and the top-p rule of inference.py:90-106 restated in plain PyTorch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,V", [(1, 1281), (64, 257216), (3, 1000)])
def test_argmax_lowest_index_ties(B, V):
    from paligemma_multimodal_system_b200 import _lib
    logits = torch.randn(B, V, device="cuda")
    logits[0, 7] = logits[0, V - 3] = 100.0  # tie -> lowest index
    out = torch.empty(B, device="cuda", dtype=torch.int32)
    _lib.check(_lib.lib().pg_argmax(logits.data_ptr(), V, out.data_ptr(), B, V, _lib.stream()), "argmax")
    torch.cuda.synchronize()
    assert torch.equal(out.long(), torch.argmax(logits, -1))
    assert out[0].item() == 7


def _kept_mask(probs, p):
    srt, idx = torch.sort(probs, dim=-1, descending=True)
    cs = torch.cumsum(srt, -1)
    drop = cs - srt > p
    keep_sorted = ~drop
    keep = torch.zeros_like(keep_sorted)
    keep.scatter_(1, idx, keep_sorted)
    return keep


@pytest.mark.parametrize("B,V,sigma", [(4, 1281, 1.0), (8, 257216, 1.0), (4, 257216, 3.0), (2, 50000, 0.1)])
def test_top_p_kept_set_and_membership(B, V, sigma):
    from paligemma_multimodal_system_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(0)
    logits = torch.randn(B, V, device="cuda", generator=g) * sigma
    temp, top_p = 0.8, 0.9
    probs = torch.softmax(logits / temp, -1)
    keep = _kept_mask(probs.double(), top_p)
    out = torch.empty(B, device="cuda", dtype=torch.int32)
    cnt = torch.empty(B, device="cuda", dtype=torch.int32)
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    for s in range(16):
        step.fill_(s)
        rc = _lib.lib().pg_sample_top_p(logits.data_ptr(), V, out.data_ptr(), cnt.data_ptr(), B, V, 1.0 / temp, top_p, 1234,
                                        step.data_ptr(), _lib.stream())
        _lib.check(rc, "top-p")
        torch.cuda.synchronize()
        # kept-set size agrees up to fp32 rounding at the boundary
        ref_cnt = keep.sum(-1)
        assert (cnt.long() - ref_cnt).abs().max().item() <= max(2, int(2e-4 * V)), (cnt, ref_cnt)
        assert keep[torch.arange(B), out.long()].all(), "sampled token outside the reference kept set"
        # without the kept-count output the rejection kernel runs (what generate() uses): same membership guarantee
        out2 = torch.full((B,), -1, device="cuda", dtype=torch.int32)
        _lib.check(_lib.lib().pg_sample_top_p(logits.data_ptr(), V, out2.data_ptr(), 0, B, V, 1.0 / temp, top_p, 1234,
                                              step.data_ptr(), _lib.stream()), "top-p (rejection)")
        torch.cuda.synchronize()
        assert keep[torch.arange(B), out2.long()].all(), "rejection sampler: token outside the reference kept set"


def test_top_p_distribution_chi2():
    from paligemma_multimodal_system_b200 import _lib
    V, draws = 64, 20000
    logits = (torch.randn(1, V, device="cuda") * 1.5).repeat(draws, 1).contiguous()
    temp, top_p = 0.8, 0.9
    probs = torch.softmax(logits[0:1].double() / temp, -1)
    keep = _kept_mask(probs, top_p)[0]
    pk = (probs[0] * keep) / (probs[0] * keep).sum()
    out = torch.empty(draws, device="cuda", dtype=torch.int32)
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    _lib.check(_lib.lib().pg_sample_top_p(logits.data_ptr(), V, out.data_ptr(), 0, draws, V, 1.0 / temp, top_p, 99,
                                          step.data_ptr(), _lib.stream()), "top-p")
    torch.cuda.synchronize()
    counts = torch.bincount(out.long(), minlength=V).double()
    assert counts[~keep].sum() == 0
    exp = pk * draws
    sel = exp > 5
    chi2 = (((counts - exp) ** 2) / exp)[sel].sum().item()
    dof = int(sel.sum().item()) - 1
    assert chi2 < dof + 6 * (2 * dof) ** 0.5, f"chi2 {chi2:.1f} for {dof} dof"


def test_top_p_degenerate_peaked_row():
    from paligemma_multimodal_system_b200 import _lib
    V = 257216
    logits = torch.zeros(2, V, device="cuda")
    logits[0, 108] = 50.0
    logits[1, 5] = 30.0
    out = torch.empty(2, device="cuda", dtype=torch.int32)
    cnt = torch.empty(2, device="cuda", dtype=torch.int32)
    _lib.check(_lib.lib().pg_sample_top_p(logits.data_ptr(), V, out.data_ptr(), cnt.data_ptr(), 2, V, 1.25, 0.9, 1, 0, _lib.stream()), "top-p")
    torch.cuda.synchronize()
    assert out.tolist() == [108, 5] and cnt.tolist() == [1, 1]
    out.fill_(-1)
    _lib.check(_lib.lib().pg_sample_top_p(logits.data_ptr(), V, out.data_ptr(), 0, 2, V, 1.25, 0.9, 1, 0, _lib.stream()), "top-p")
    torch.cuda.synchronize()
    assert out.tolist() == [108, 5]


@pytest.mark.parametrize("sigma,top_p", [(1.5, 0.9), (0.3, 0.5), (4.0, 0.95)])
def test_top_p_rejection_sampler_chi2_large_vocab(sigma, top_p):
    """The rejection kernel over a cluster of two CTAs (V >= 65536): empirical distribution of 4096 draws of ONE row
    against the renormalised kept set, aggregated into 32 probability-ordered buckets."""
    from paligemma_multimodal_system_b200 import _lib
    V, draws, temp = 70000, 4096, 0.8
    g = torch.Generator(device="cuda").manual_seed(3)
    row = torch.randn(1, V, device="cuda", generator=g) * sigma
    logits = row.repeat(draws, 1).contiguous()
    probs = torch.softmax(row.double() / temp, -1)
    keep = _kept_mask(probs, top_p)[0]
    pk = (probs[0] * keep) / (probs[0] * keep).sum()
    out = torch.empty(draws, device="cuda", dtype=torch.int32)
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    _lib.check(_lib.lib().pg_sample_top_p(logits.data_ptr(), V, out.data_ptr(), 0, draws, V, 1.0 / temp, top_p, 7,
                                          step.data_ptr(), _lib.stream()), "top-p")
    torch.cuda.synchronize()
    assert keep[out.long()].all()
    order = torch.argsort(pk, descending=True)
    cdf = torch.cumsum(pk[order], 0)
    bucket_of = torch.empty(V, dtype=torch.long, device="cuda")
    bucket_of[order] = torch.clamp((cdf * 32).long(), max=31)
    exp = torch.zeros(32, dtype=torch.double, device="cuda").index_add_(0, bucket_of, pk) * draws
    cnt = torch.bincount(bucket_of[out.long()], minlength=32).double()
    sel = exp > 5
    chi2 = (((cnt - exp) ** 2) / exp)[sel].sum().item()
    dof = int(sel.sum().item()) - 1
    assert chi2 < dof + 6 * (2 * dof) ** 0.5, f"chi2 {chi2:.1f} for {dof} dof"


# ---------------------------------------------------------------------------------------------------------------------
# samplers fed by the lm_head epilogue's segment statistics (pg_gemm_bf16_fused stats -> pg_*_stats)
# ---------------------------------------------------------------------------------------------------------------------
def _torch_stats(logits, inv_temp, nseg=None):
    """(max, sum exp2((x - max) * inv_temp * log2 e)) of every 32-token segment, as the GEMM epilogue defines them; laid out
    [segment, row, 2] like the epilogue writes them."""
    B, V = logits.shape
    n = (V + 31) // 32
    pad = torch.full((B, n * 32), float("-inf"), device=logits.device)
    pad[:, :V] = logits
    seg = pad.view(B, n, 32)
    m = seg.max(-1).values
    s = torch.exp2((seg - m[..., None]) * (inv_temp * 1.4426950408889634)).sum(-1)
    out = torch.zeros(nseg or n, B, 2, device=logits.device)
    out[:, :, 0] = float("-inf")
    out[:n, :, 0], out[:n, :, 1] = m.t(), s.t()
    return out.contiguous()


@pytest.mark.parametrize("T,V,K", [(64, 257216, 2048), (5, 1281, 256), (33, 4000, 512), (128, 1281, 256), (16, 70001, 256)])
def test_lm_head_epilogue_statistics(T, V, K):
    """The fused lm_head (logits + bias AND per-segment max / sum-exp2) against torch on its own fp32 logits."""
    from paligemma_multimodal_system_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(1)
    Vp = (V + 7) // 8 * 8
    x = (torch.randn(T, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(V, K, device="cuda", generator=g) * 0.08).bfloat16()
    bias = torch.randn(V, device="cuda", generator=g)
    inv_temp = 1.25
    nseg = 4 * ((V + 127) // 128)
    logits = torch.full((T, V), float("nan"), device="cuda")
    stats = torch.full((nseg, T, 2), float("nan"), device="cuda")
    _lib.gemm_fused(x, w, logits, mode=_lib.EPI_F32, bias=bias, stats=stats, inv_temperature=inv_temp)
    plain = torch.empty(T, V, device="cuda")
    _lib.gemm(x, w, plain, mode=_lib.EPI_F32, bias=bias, swap=1)
    torch.cuda.synchronize()
    assert torch.equal(logits, plain), "the statistics epilogue must write the very same logits"
    ref = _torch_stats(logits, inv_temp, nseg)
    assert torch.equal(stats[..., 0], ref[..., 0]), "segment maxima are exact"
    assert (stats[..., 1] - ref[..., 1]).abs().max().item() <= 2e-5 * 32  # (2^-24 fixed-point sum of <= 32 terms in [0, 1])
    assert not torch.isnan(stats).any()


@pytest.mark.parametrize("B,V", [(1, 1281), (64, 257216), (3, 1000), (2, 33)])
def test_argmax_from_statistics_lowest_index_ties(B, V):
    from paligemma_multimodal_system_b200 import _lib
    logits = torch.randn(B, V, device="cuda")
    logits[0, 7] = logits[0, V - 3] = 100.0          # tie across segments -> lowest index
    if B > 1:
        logits[1, 20] = logits[1, 25] = 50.0          # tie inside one segment
    stats = _torch_stats(logits, 1.0)
    out = torch.full((B,), -1, device="cuda", dtype=torch.int32)
    _lib.check(_lib.lib().pg_argmax_stats(logits.data_ptr(), V, stats.data_ptr(), stats.shape[1], out.data_ptr(), B, V, _lib.stream()), "argmax")
    torch.cuda.synchronize()
    assert torch.equal(out.long(), torch.argmax(logits, -1))
    assert out[0].item() == 7


@pytest.mark.parametrize("B,V,sigma", [(4, 1281, 1.0), (8, 257216, 1.0), (4, 257216, 3.0), (2, 50000, 0.1), (64, 257216, 2.0)])
def test_top_p_from_statistics_membership(B, V, sigma):
    from paligemma_multimodal_system_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(0)
    logits = torch.randn(B, V, device="cuda", generator=g) * sigma
    temp, top_p = 0.8, 0.9
    keep = _kept_mask(torch.softmax(logits.double() / temp, -1), top_p)
    stats = _torch_stats(logits, 1.0 / temp)
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    seed_dev = torch.tensor([77], device="cuda", dtype=torch.int64)
    seen = set()
    for s in range(16):
        step.fill_(s)
        out = torch.full((B,), -1, device="cuda", dtype=torch.int32)
        _lib.check(_lib.lib().pg_sample_top_p_stats(logits.data_ptr(), V, stats.data_ptr(), stats.shape[1], out.data_ptr(), B, V,
                                                    1.0 / temp, top_p, 1234, seed_dev.data_ptr(), step.data_ptr(), _lib.stream()), "top-p stats")
        torch.cuda.synchronize()
        assert keep[torch.arange(B), out.long()].all(), "token outside the reference kept set"
        out2 = torch.full((B,), -1, device="cuda", dtype=torch.int32)
        _lib.check(_lib.lib().pg_sample_top_p_stats(logits.data_ptr(), V, stats.data_ptr(), stats.shape[1], out2.data_ptr(), B, V,
                                                    1.0 / temp, top_p, 77, 0, step.data_ptr(), _lib.stream()), "top-p stats")
        torch.cuda.synchronize()
        assert torch.equal(out, out2), "the device seed overrides the by-value seed; same seed -> same draw"
        seen.add(tuple(out.tolist()))
    assert len(seen) > 1


@pytest.mark.parametrize("V,sigma,top_p", [(64, 1.5, 0.9), (70000, 1.5, 0.9), (70000, 0.3, 0.5), (257216, 4.0, 0.95)])
def test_top_p_from_statistics_chi2(V, sigma, top_p):
    """Empirical distribution of 4096 draws of one row against the renormalised kept set (32 probability-ordered buckets)."""
    from paligemma_multimodal_system_b200 import _lib
    draws, temp = 4096 if V > 64 else 20000, 0.8
    if V > 100000:
        draws = 1024
    g = torch.Generator(device="cuda").manual_seed(3)
    row = torch.randn(1, V, device="cuda", generator=g) * sigma
    logits = row.repeat(draws, 1).contiguous()
    probs = torch.softmax(row.double() / temp, -1)
    keep = _kept_mask(probs, top_p)[0]
    pk = (probs[0] * keep) / (probs[0] * keep).sum()
    stats = _torch_stats(row, 1.0 / temp).repeat(1, draws, 1).contiguous()
    out = torch.empty(draws, device="cuda", dtype=torch.int32)
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    _lib.check(_lib.lib().pg_sample_top_p_stats(logits.data_ptr(), V, stats.data_ptr(), stats.shape[1], out.data_ptr(), draws, V,
                                                1.0 / temp, top_p, 7, 0, step.data_ptr(), _lib.stream()), "top-p stats")
    torch.cuda.synchronize()
    assert keep[out.long()].all()
    nb = 32 if V > 64 else 16
    order = torch.argsort(pk, descending=True)
    cdf = torch.cumsum(pk[order], 0)
    bucket_of = torch.empty(V, dtype=torch.long, device="cuda")
    bucket_of[order] = torch.clamp((cdf * nb).long(), max=nb - 1)
    exp = torch.zeros(nb, dtype=torch.double, device="cuda").index_add_(0, bucket_of, pk) * draws
    cnt = torch.bincount(bucket_of[out.long()], minlength=nb).double()
    sel = exp > 5
    chi2 = (((cnt - exp) ** 2) / exp)[sel].sum().item()
    dof = int(sel.sum().item()) - 1
    assert chi2 < dof + 6 * (2 * dof) ** 0.5, f"chi2 {chi2:.1f} for {dof} dof"


def test_top_p_tiny_threshold_falls_back_to_the_most_probable_token():
    """top_p far below the largest probability on a flat row: every candidate is rejected (acceptance ~ top_p per draw); both
    rejection samplers then emit the row argmax, which is always inside the kept set -- never a stale / unwritten token."""
    from paligemma_multimodal_system_b200 import _lib
    B, V = 3, 70000
    g = torch.Generator(device="cuda").manual_seed(5)
    logits = torch.randn(B, V, device="cuda", generator=g) * 0.05
    logits[:, 4321] += 0.01
    am = torch.argmax(logits, -1)
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    stats = _torch_stats(logits, 1.0)
    for fn in ("stats", "plain"):
        out = torch.full((B,), -1, device="cuda", dtype=torch.int32)
        if fn == "stats":
            rc = _lib.lib().pg_sample_top_p_stats(logits.data_ptr(), V, stats.data_ptr(), stats.shape[1], out.data_ptr(), B, V, 1.0, 1e-7, 3, 0,
                                                  step.data_ptr(), _lib.stream())
        else:
            rc = _lib.lib().pg_sample_top_p(logits.data_ptr(), V, out.data_ptr(), 0, B, V, 1.0, 1e-7, 3, step.data_ptr(), _lib.stream())
        _lib.check(rc, fn)
        torch.cuda.synchronize()
        assert torch.equal(out.long(), am), fn
