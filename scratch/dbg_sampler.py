import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
V, draws = 64, 20000
temp, top_p = 0.8, 0.9
for seed in range(40):
    torch.manual_seed(seed)
    logits = (torch.randn(1, V, device="cuda") * 1.5).repeat(draws, 1).contiguous()
    probs = torch.softmax(logits[0:1].double() / temp, -1)
    srt, idx = torch.sort(probs, dim=-1, descending=True)
    cs = torch.cumsum(srt, -1)
    keep_sorted = ~(cs - srt > top_p)
    keep = torch.zeros_like(keep_sorted).scatter_(1, idx, keep_sorted)[0]
    pk = (probs[0] * keep) / (probs[0] * keep).sum()
    out = torch.empty(draws, device="cuda", dtype=torch.int32)
    cnt = torch.empty(draws, device="cuda", dtype=torch.int32)
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    _lib.check(_lib.lib().pg_sample_top_p(logits.data_ptr(), V, out.data_ptr(), cnt.data_ptr(), draws, V, 1.0 / temp, top_p, 99, step.data_ptr(), _lib.stream()), "x")
    torch.cuda.synchronize()
    counts = torch.bincount(out.long(), minlength=V).double()
    exp = pk * draws
    sel = exp > 5
    chi2 = (((counts - exp) ** 2) / exp)[sel].sum().item()
    bad = chi2 > int(sel.sum()) + 40
    print(seed, "ref kept", int(keep.sum()), "kernel kept", cnt[0].item(), "chi2 %.1f" % chi2, "outside", int(counts[~keep].sum()), "BAD" if bad else "")
    if bad:
        order = torch.argsort(pk, descending=True)
        for i in order[: int(keep.sum()) + 1].tolist():
            print("   ", i, "logit=%.6f p=%.4f exp=%.1f got=%d keep=%d" % (logits[0, i], pk[i], pk[i] * draws, counts[i], keep[i]))
