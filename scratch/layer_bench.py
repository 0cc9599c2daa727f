"""Steady-state timing of the NEW decode layer chain (cluster split-K GEMMs, RMSNorm folded into the GEMM epilogues)."""
import os, sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
B, D, F, Hq, Hkv, dh, V, NL = int(os.environ.get("MB_B", 64)), 2048, 16384, 8, 1, 256, 257216, 18
W = (Hq + 2 * Hkv) * dh
dev = "cuda"
def rnd(*s): return (torch.randn(*s, device=dev) * 0.02).bfloat16()
qkv_w = [rnd(W, D) for _ in range(NL)]; o_w = [rnd(D, D) for _ in range(NL)]
gu_w = [rnd(2 * F, D) for _ in range(NL)]; down_w = [rnd(D, F) for _ in range(NL)]
hb = rnd(B, D); att = rnd(B, Hq * dh); mid = rnd(B, F)
h = torch.randn(B, D, device=dev); qkv = torch.zeros(B, W, device=dev)
ss = torch.rand(2 * NL + 1, B, device=dev) * D + 1
ln_w = torch.zeros(D, device=dev)
PAGE = 64; kvlen = 324; max_pages = 7
k_pages = [rnd(B * max_pages, PAGE, dh) for _ in range(NL)]; v_pages = [rnd(B * max_pages, PAGE, dh) for _ in range(NL)]
table = torch.arange(B * max_pages, device=dev, dtype=torch.int32).view(B, max_pages).contiguous()
kvl = torch.full((B,), kvlen, device=dev, dtype=torch.int32); posd = torch.full((B,), kvlen, device=dev, dtype=torch.int32)
inv_freq = (1.0 / (10000.0 ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))).to(dev)
def attn(i):
    _lib.check(L.pg_attention_decode_fused(qkv.data_ptr(), posd.data_ptr(), kvl.data_ptr(), inv_freq.data_ptr(), k_pages[i].data_ptr(), v_pages[i].data_ptr(), table.data_ptr(), att.data_ptr(), B, Hq, Hkv, dh, PAGE, B * max_pages, max_pages, 1.0 / 16, _lib.stream()), "attn")

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
    print(f"{name:44s} {us:8.2f} us   {nbytes / us / 1e3:8.1f} GB/s  ideal {nbytes / 6550.7e3:6.2f} us", flush=True)

for S in (4, 8, 16):
    graph_time(f"qkv cluster {S}", lambda i: _lib.gemm_decode(hb, qkv_w[i], qkv, mode=0, cluster_k=S, ss_in=ss[0], norm_dim=D), W * D * 2)
for S in (4, 8, 16):
    graph_time(f"o cluster {S} (+resid+norm operands)", lambda i: _lib.gemm_decode(att, o_w[i], h, mode=1, cluster_k=S, hb=hb, norm_w=ln_w, ss_out=ss[1]), D * D * 2)
for S in (8, 16):
    graph_time(f"down cluster {S} (+resid+norm operands)", lambda i: _lib.gemm_decode(mid, down_w[i], h, mode=1, cluster_k=S, hb=hb, norm_w=ln_w, ss_out=ss[1]), D * F * 2)
graph_time("gate-up geglu colnorm", lambda i: _lib.gemm_colnorm(hb, gu_w[i], mid, mode=_lib.EPI_GEGLU, ss_in=ss[1], norm_dim=D), 2 * F * D * 2)
graph_time("attention fused (kv 324)", attn, B * kvlen * dh * 2 * 2)
nbytes = (W * D + D * D + 3 * F * D) * 2 + B * kvlen * dh * 4
for sq, so, sd in ((8, 8, 8), (8, 8, 16), (8, 16, 16), (4, 8, 16), (8, 4, 16)):
    def full_layer(i):
        _lib.gemm_decode(hb, qkv_w[i], qkv, mode=0, cluster_k=sq, ss_in=ss[2 * i], norm_dim=D)
        attn(i)
        _lib.gemm_decode(att, o_w[i], h, mode=1, cluster_k=so, hb=hb, norm_w=ln_w, ss_out=ss[2 * i + 1])
        _lib.gemm_colnorm(hb, gu_w[i], mid, mode=_lib.EPI_GEGLU, ss_in=ss[2 * i + 1], norm_dim=D)
        _lib.gemm_decode(mid, down_w[i], h, mode=1, cluster_k=sd, hb=hb, norm_w=ln_w, ss_out=ss[2 * i + 2])
    graph_time(f"FULL layer S=({sq},{so},{sd})", full_layer, nbytes)
def no_attn(i):
    _lib.gemm_decode(hb, qkv_w[i], qkv, mode=0, cluster_k=8, ss_in=ss[2 * i], norm_dim=D)
    _lib.gemm_decode(att, o_w[i], h, mode=1, cluster_k=8, hb=hb, norm_w=ln_w, ss_out=ss[2 * i + 1])
    _lib.gemm_colnorm(hb, gu_w[i], mid, mode=_lib.EPI_GEGLU, ss_in=ss[2 * i + 1], norm_dim=D)
    _lib.gemm_decode(mid, down_w[i], h, mode=1, cluster_k=16, hb=hb, norm_w=ln_w, ss_out=ss[2 * i + 2])
graph_time("layer without attention", no_attn, nbytes)
