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
