import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
B, D, F, NL = 64, 2048, 16384, 18
def rnd(*s): return (torch.randn(*s, device="cuda") * 0.02).bfloat16()
o_w = [rnd(D, D) for _ in range(NL)]; down_w = [rnd(D, F) for _ in range(NL)]; gu_w=[rnd(2*F, D) for _ in range(NL)]
att = rnd(B, D); mid = rnd(B, F); h = torch.randn(B, D, device="cuda"); midout = torch.empty(B, F, device="cuda", dtype=torch.bfloat16)
tr = torch.zeros(8 * 64, device="cuda", dtype=torch.int64)
def run(name, fn):
    L.pg_debug_set_gemm_trace(0)
    for i in range(NL): fn(i)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    L.pg_debug_set_gemm_trace(tr.data_ptr())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(NL): fn(i)
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    t = tr.cpu().numpy().astype("float64").reshape(64, 8)
    for k in (8, 9, 10):
        d = (t[k, 1:8] - t[k, 0]) / 1.9e3
        nxt = (t[k + 1, 0] - t[k, 0]) / 1.9e3
        print(f"{name:14s} launch {k}: prefetch-issued {d[0]:5.2f} | prev-done {d[1]:5.2f} | acc ready {d[2]:5.2f} | epilogue issued {d[3]:5.2f} | all done t0 {d[4]:5.2f} t64 {d[5]:5.2f} t32 {d[6]:5.2f} | next kernel entry {nxt:5.2f}")
run("o split9", lambda i: _lib.gemm(att, o_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=9))
run("down split18", lambda i: _lib.gemm(mid, down_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=18))
run("gate-up", lambda i: _lib.gemm(att, gu_w[i], midout, mode=_lib.EPI_GEGLU, swap=1))
