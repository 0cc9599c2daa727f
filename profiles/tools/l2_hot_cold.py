"""Does the swap-AB decode GEMM run faster when its weights are L2-resident?  (decides whether any L2 weight prefetch can pay)
gate||up slices of 16..128 MB: cold (18 matrices in rotation), hot (same matrix every launch), and cold-but-touched
(a torch reduction reads the matrix right before the GEMM)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from paligemma_multimodal_system_b200 import _lib
B, D = int(os.environ.get("MB_B", 64)), 2048
dev = "cuda"
hn = (torch.randn(B, D, device=dev) * 0.02).bfloat16()

def timeit(fn, n, reps=5):
    for i in range(n): fn(i)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n): fn(i)
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): g.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (reps * n))
    return best

for rows in (4096, 8192, 16384, 24576, 32768):
    mb = rows * D * 2 / 1e6
    ws = [(torch.randn(rows, D, device=dev) * 0.02).bfloat16() for _ in range(18)]
    out = torch.empty(B, rows // 2, device=dev, dtype=torch.bfloat16)
    sink = torch.empty(18, device=dev)
    cold = timeit(lambda i: _lib.gemm(hn, ws[i], out, mode=_lib.EPI_GEGLU, swap=1), 18)
    hot = timeit(lambda i: _lib.gemm(hn, ws[0], out, mode=_lib.EPI_GEGLU, swap=1), 18)
    def touched(i):
        torch.sum(ws[i].view(torch.int16), dtype=torch.int32, out=sink[i].view(torch.int32)) if False else sink[i:i+1].copy_(ws[i].view(torch.int32).sum().float().view(1))
        _lib.gemm(hn, ws[i], out, mode=_lib.EPI_GEGLU, swap=1)
    def touch_only(i):
        sink[i:i+1].copy_(ws[i].view(torch.int32).sum().float().view(1))
    t_both = timeit(touched, 18)
    t_touch = timeit(touch_only, 18)
    print(f"{mb:7.1f} MB: cold {cold:6.2f} us ({mb*1e3/cold/1e3:5.2f} TB/s) | hot {hot:6.2f} us ({mb*1e3/hot/1e3:5.2f} TB/s) | touch+gemm {t_both:6.2f}, touch alone {t_touch:6.2f} -> gemm after touch {t_both - t_touch:6.2f} us", flush=True)
    del ws
