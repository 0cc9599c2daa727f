import sys, torch, numpy as np
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
B, D, F = 64, 2048, 16384
def rnd(*s): return (torch.randn(*s, device="cuda") * 0.02).bfloat16()
down_w = rnd(D, F); o_w = rnd(D, D); big_w = rnd(296 * 128 // 8, D)
att = rnd(B, D); mid = rnd(B, F); h = torch.randn(B, D, device="cuda"); hb = rnd(B, D)
ss = torch.ones(B, device="cuda"); ln_w = torch.zeros(D, device="cuda")
out_big = torch.zeros(B, big_w.shape[0], device="cuda")
tr = torch.zeros(3 * 1024, device="cuda", dtype=torch.int64)
def report(name, ncta):
    torch.cuda.synchronize()
    t = tr.cpu().numpy().reshape(-1, 3)[:ncta]
    t0 = t[:, 1].min()
    sm = t[:, 0]
    per_sm = np.bincount(sm.astype(int), minlength=148)
    # overlapping CTAs on the same SM?
    overlap = 0
    for s in np.unique(sm):
        iv = sorted((a, b) for a, b in t[sm == s][:, 1:3])
        for i in range(len(iv) - 1):
            if iv[i + 1][0] < iv[i][1]: overlap += 1
    print(f"{name}: {ncta} CTAs on {len(np.unique(sm))} SMs, max CTAs/SM {per_sm.max()}, co-resident pairs {overlap}, "
          f"start spread {(t[:,1].max()-t0)/1e3:.2f} us, end {(t[:,2].max()-t0)/1e3:.2f} us, first end {(t[:,2].min()-t0)/1e3:.2f} us")
L.pg_set_pdl(0)
for S in (1, 2, 4, 8, 16):
    for w, x, nm in ((o_w, att, "o"), (down_w, mid, "down")):
        _lib.gemm_decode(x, w, h, mode=1, cluster_k=S, hb=hb, norm_w=ln_w, ss_out=ss)
        torch.cuda.synchronize()
        tr.zero_()
        L.pg_debug_set_decode_gemm_cta_trace(tr.data_ptr())
        _lib.gemm_decode(x, w, h, mode=1, cluster_k=S, hb=hb, norm_w=ln_w, ss_out=ss)
        L.pg_debug_set_decode_gemm_cta_trace(0)
        report(f"{nm} S={S}", 16 * S)
# 296 independent CTAs (S=1, 296 tiles)
_lib.gemm_decode(hb, big_w, out_big, mode=0, cluster_k=1)
tr.zero_(); L.pg_debug_set_decode_gemm_cta_trace(tr.data_ptr())
_lib.gemm_decode(hb, big_w, out_big, mode=0, cluster_k=1)
L.pg_debug_set_decode_gemm_cta_trace(0)
report("296 tiles S=1", big_w.shape[0] // 128)
