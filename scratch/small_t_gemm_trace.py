"""Where do ~10 us per small token-major GEMM go at T = 256 (B = 1 prefill)?  Graph of the four SigLIP projections per layer,
CTA-0 stamps (us after kernel entry) + distance to the next kernel's entry."""
import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
T, Dv, Fv, NL = 256, 1152, 4304, 6
def rnd(*s): return (torch.randn(*s, device="cuda") * 0.05).bfloat16()
qkv_w = [rnd(3 * Dv, Dv) for _ in range(NL)]; out_w = [rnd(Dv, Dv) for _ in range(NL)]
fc1_w = [rnd(Fv, Dv) for _ in range(NL)]; fc2_w = [rnd(Dv, Fv) for _ in range(NL)]
x = rnd(T, Dv); qkv = torch.empty(T, 3 * Dv, device="cuda", dtype=torch.bfloat16); att = rnd(T, Dv)
h = torch.randn(T, Dv, device="cuda"); mid = torch.empty(T, Fv, device="cuda", dtype=torch.bfloat16)
b3, b1, bf = torch.randn(3 * Dv, device="cuda"), torch.randn(Dv, device="cuda"), torch.randn(Fv, device="cuda")
tr = torch.zeros(8 * 64, device="cuda", dtype=torch.int64)
def layer(i):
    _lib.gemm(x, qkv_w[i], qkv, mode=_lib.EPI_BF16, bias=b3, swap=0)
    _lib.gemm_residual(att, out_w[i], h, bias=b1)
    _lib.gemm(x, fc1_w[i], mid, mode=_lib.EPI_BF16, bias=bf, act_gelu=True, swap=0)
    _lib.gemm_residual(mid, fc2_w[i], h, bias=b1)
L.pg_debug_set_gemm_trace(0)
for i in range(NL): layer(i)
torch.cuda.synchronize()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
L.pg_debug_set_gemm_trace(tr.data_ptr())
with torch.cuda.stream(s):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(NL): layer(i)
torch.cuda.current_stream().wait_stream(s)
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"graph of {4 * NL} GEMMs: {e0.elapsed_time(e1) * 1e3:.1f} us -> {e0.elapsed_time(e1) * 1e3 / (4 * NL):.1f} us per GEMM")
t = tr.cpu().numpy().astype("float64").reshape(64, 8)
names = ["qkv", "out", "fc1", "fc2"]
for k in range(8, 16):
    d = (t[k, 1:8] - t[k, 0]) / 1.9e3
    nxt = (t[k + 1, 0] - t[k, 0]) / 1.9e3
    print(f"{names[k % 4]:4s}: prefetch-issued {d[0]:5.2f} | prev-done {d[1]:5.2f} | acc ready {d[2]:5.2f} | epilogue issued {d[3]:5.2f} | done t0 {d[4]:5.2f} t64 {d[5]:5.2f} t32 {d[6]:5.2f} | next entry {nxt:5.2f}")
