"""Is the weight-streaming GEMM faster when its weights are L2 resident?  Same weights every launch vs rotating."""
import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
B, D, F, NL = 64, 2048, 16384, 18
dev = "cuda"
def rnd(*s): return (torch.randn(*s, device=dev) * 0.02).bfloat16()
down_w = [rnd(D, F) for _ in range(NL)]
gu_half = [rnd(F, D) for _ in range(NL)]   # 64 MB "gate||up" with F/2
gu_q = [rnd(F // 2, D) for _ in range(NL)]   # 32 MB
hn = rnd(B, D); mid = rnd(B, F); h = torch.zeros(B, D, device=dev)
out = torch.empty(B, F, device=dev, dtype=torch.bfloat16)

def graph_time(name, fn, nbytes, reps=5):
    for i in range(NL): fn(i)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(NL): fn(i)
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * NL)
    print(f"{name:50s} {us:8.2f} us  {nbytes / us / 1e3:8.1f} GB/s", flush=True)

for pdl in (1, 0):
    L.pg_set_pdl(pdl)
    print("PDL", pdl)
    graph_time("down 67MB rotating, split18", lambda i: _lib.gemm(mid, down_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=18), D * F * 2)
    graph_time("down 67MB SAME weights, split18", lambda i: _lib.gemm(mid, down_w[0], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=18), D * F * 2)
    graph_time("geglu 64MB rotating", lambda i: _lib.gemm(hn, gu_half[i], out, mode=_lib.EPI_GEGLU, swap=1), D * F * 2)
    graph_time("geglu 64MB SAME weights", lambda i: _lib.gemm(hn, gu_half[0], out, mode=_lib.EPI_GEGLU, swap=1), D * F * 2)
    graph_time("geglu 32MB rotating", lambda i: _lib.gemm(hn, gu_q[i], out, mode=_lib.EPI_GEGLU, swap=1), D * F)
    graph_time("geglu 32MB SAME weights", lambda i: _lib.gemm(hn, gu_q[0], out, mode=_lib.EPI_GEGLU, swap=1), D * F)
    # prefetch then run: side effect of the LSU prefetch
    def pf_then(i):
        _lib.check(L.pg_prefetch_l2(gu_half[i].data_ptr(), gu_half[i].numel() * 2, 0, 148, _lib.stream()), "pf")
    graph_time("prefetch kernel alone 64MB mode0", pf_then, D * F * 2)
    def pf_then1(i):
        _lib.check(L.pg_prefetch_l2(gu_half[i].data_ptr(), gu_half[i].numel() * 2, 1, 148, _lib.stream()), "pf")
    graph_time("prefetch kernel alone 64MB mode1", pf_then1, D * F * 2)
    def pf_gemm(i):
        _lib.check(L.pg_prefetch_l2(gu_half[i].data_ptr(), gu_half[i].numel() * 2, 1, 148, _lib.stream()), "pf")
        _lib.gemm(hn, gu_half[i], out, mode=_lib.EPI_GEGLU, swap=1)
    graph_time("prefetch mode1 + geglu 64MB (serial)", pf_gemm, D * F * 2)
