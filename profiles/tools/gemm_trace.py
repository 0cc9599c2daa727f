"""clock64 stamps of CTA 0 of the decode (swap-AB) GEMMs inside a CUDA-graph chain: where a launch spends its time."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
B, D, F = 64, 2048, 16384
dev = "cuda"
def rnd(*s): return (torch.randn(*s, device=dev) * 0.02).bfloat16()
hn = rnd(B, D); mid = rnd(B, F)
tr = torch.zeros(8 * 64, device=dev, dtype=torch.int64)
names = ["prefetch issued", "wait returned", "acc ready", "epilogue issued", "all done w0", "w2", "w1"]

def run(name, fn, n=18):
    for i in range(n): fn(i)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    L.pg_debug_set_gemm_trace(tr.data_ptr())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n): fn(i)
    L.pg_debug_set_gemm_trace(0)
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    t = tr.cpu().numpy().astype("float64").reshape(64, 8)
    print(f"{name}: {e0.elapsed_time(e1) * 1e3 / n:.2f} us per launch")
    for k in (8, 9, 10):
        d = (t[k, 1:8] - t[k, 0]) / 1.9e3
        print(f"   launch {k}: " + " | ".join(f"{nm} {v:6.2f}" for nm, v in zip(names, d)) + f" | next entry {(t[k + 1, 0] - t[k, 0]) / 1.9e3:6.2f}")

for rows in (4096, 32768):
    ws = [rnd(rows, D) for _ in range(18)]
    out = torch.empty(B, rows // 2, device=dev, dtype=torch.bfloat16)
    run(f"gate||up GEGLU {rows} rows hot ", lambda i: _lib.gemm(hn, ws[0], out, mode=_lib.EPI_GEGLU, swap=1))
    run(f"gate||up GEGLU {rows} rows cold", lambda i: _lib.gemm(hn, ws[i], out, mode=_lib.EPI_GEGLU, swap=1))
    outb = torch.empty(B, rows, device=dev, dtype=torch.bfloat16)
    run(f"plain bf16 {rows} rows hot ", lambda i: _lib.gemm(hn, ws[0], outb, mode=_lib.EPI_BF16, swap=1))
    del ws
dw = [rnd(D, F) for _ in range(18)]
h = torch.zeros(B, D, device=dev)
for sp in (18, 9):
    run(f"down split-K {sp} cold", lambda i: _lib.gemm(mid, dw[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sp))
qw = [rnd(2560, D) for _ in range(18)]
qkv = torch.zeros(B, 2560, device=dev)
run("qkv split-K 7 cold", lambda i: _lib.gemm(hn, qw[i], qkv, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=7))
