import os, sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
B, D, F, NL = 64, 2048, 16384, 18
def rnd(*s): return (torch.randn(*s, device="cuda") * 0.02).bfloat16()
o_w = [rnd(D, D) for _ in range(NL)]; down_w = [rnd(D, F) for _ in range(NL)]; qkv_w = [rnd(2560, D) for _ in range(NL)]
att = rnd(B, D); mid = rnd(B, F); h = torch.randn(B, D, device="cuda"); hb = rnd(B, D); qkv = torch.zeros(B, 2560, device="cuda")
ss = torch.ones(B, device="cuda"); ln_w = torch.zeros(D, device="cuda")
tr = torch.zeros(8 * 64, device="cuda", dtype=torch.int64)
def run(name, fn):
    L.pg_debug_set_decode_gemm_trace(0)
    for i in range(NL): fn(i)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    L.pg_debug_set_decode_gemm_trace(tr.data_ptr())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(NL): fn(i)
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    t = tr.cpu().numpy().astype("float64").reshape(64, 8)
    print(f"{name}: {e0.elapsed_time(e1) * 1e3 / NL:.2f} us per launch")
    for k in (8, 9):
        d = (t[k, 1:8] - t[k, 0]) / 1.9e3
        nxt = (t[k + 1, 0] - t[k, 0]) / 1.9e3
        print(f"   launch {k}: weights issued {d[0]:5.2f} | prev done {d[1]:5.2f} | acc ready {d[2]:5.2f} | cluster bar1 {d[3]:5.2f} | scattered {d[4]:5.2f} | cluster bar2 {d[5]:5.2f} | finalized {d[6]:5.2f} | next entry {nxt:5.2f}")
    L.pg_debug_set_decode_gemm_trace(0)
for pdl in (1, 0):
    L.pg_set_pdl(pdl); print("PDL", pdl)
    for S in (2, 4, 8):
        run(f"o cluster {S}", lambda i: _lib.gemm_decode(att, o_w[i], h, mode=1, cluster_k=S, hb=hb, norm_w=ln_w, ss_out=ss))
    run("qkv cluster 8", lambda i: _lib.gemm_decode(hb, qkv_w[i], qkv, mode=0, cluster_k=8, ss_in=ss, norm_dim=D))
    run("qkv cluster 4", lambda i: _lib.gemm_decode(hb, qkv_w[i], qkv, mode=0, cluster_k=4, ss_in=ss, norm_dim=D))
    run("down cluster 4", lambda i: _lib.gemm_decode(mid, down_w[i], h, mode=1, cluster_k=4, hb=hb, norm_w=ln_w, ss_out=ss))
    run("down cluster 16", lambda i: _lib.gemm_decode(mid, down_w[i], h, mode=1, cluster_k=16, hb=hb, norm_w=ln_w, ss_out=ss))
    run("down cluster 8", lambda i: _lib.gemm_decode(mid, down_w[i], h, mode=1, cluster_k=8, hb=hb, norm_w=ln_w, ss_out=ss))
