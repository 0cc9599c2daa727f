import sys, torch
sys.path.insert(0, '.')  # run from the repo root
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
B, Hq, Hkv, dh, NL, PAGE, kvlen, max_pages = 64, 8, 1, 256, 18, 64, 324, 7
W = (Hq + 2) * dh
def rnd(*s): return (torch.randn(*s, device="cuda") * 0.02).bfloat16()
k_pages=[rnd(B*max_pages, PAGE, dh) for _ in range(NL)]; v_pages=[rnd(B*max_pages, PAGE, dh) for _ in range(NL)]
table=torch.arange(B*max_pages, device="cuda", dtype=torch.int32).view(B,max_pages).contiguous()
kvl=torch.full((B,), kvlen, device="cuda", dtype=torch.int32); posd=kvl.clone()
inv_freq=(1.0/(10000.0**(torch.arange(0,dh,2,dtype=torch.int64).float()/dh))).cuda()
qkvf=torch.randn(B, W, device="cuda")*0.5
out=torch.empty(B, Hq*dh, device="cuda", dtype=torch.bfloat16)
tr = torch.zeros(8 * 64 + 256, device="cuda", dtype=torch.int64)
def attn(i):
    _lib.check(L.pg_attention_decode_fused(qkvf.data_ptr(), posd.data_ptr(), kvl.data_ptr(), inv_freq.data_ptr(), k_pages[i].data_ptr(), v_pages[i].data_ptr(), table.data_ptr(), out.data_ptr(), B, Hq, Hkv, dh, PAGE, B*max_pages, max_pages, 1.0/16, _lib.stream()), "attn")
for i in range(NL): attn(i)
torch.cuda.synchronize()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
L.pg_debug_set_attn_trace(tr.data_ptr())
with torch.cuda.stream(s):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(NL): attn(i)
torch.cuda.current_stream().wait_stream(s)
for _ in range(3): g.replay()
torch.cuda.synchronize()
t = tr.cpu().numpy().astype("float64")[:512].reshape(64, 8)
import os
names = ["wait-returned", "Q staged", "pages landed", "rounds done", "cluster-sync1", "end", "-"]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) * 1e3 / NL:.2f} us per launch")
for k in (8, 9):
    d = (t[k, 1:8] - t[k, 0]) / 1.9e3
    d1 = (t[k + 32, 1:8] - t[k, 0]) / 1.9e3
    print(f"   rank 1 (entry {(t[k+32,0]-t[k,0])/1.9e3:5.2f}): " + " | ".join(f"{n} {v:5.2f}" for n, v in zip(names, d1)))
    print(f"launch {k}: " + " | ".join(f"{n} {v:5.2f}" for n, v in zip(names, d)) + f" | next entry {(t[k+1,0]-t[k,0])/1.9e3:5.2f}")
