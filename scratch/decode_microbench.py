"""Steady-state timing of each decode-step kernel (3B shapes, B=64), rotating over 18 layers' weights (cold L2)."""
import math, os, sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
B, D, F, Hq, Hkv, dh, V, NL = int(os.environ.get("MB_B", 64)), 2048, 16384, 8, 1, 256, 257216, 18
W = (Hq + 2 * Hkv) * dh
dev = "cuda"
def rnd(*s): return (torch.randn(*s, device=dev) * 0.02).bfloat16()
qkv_w = [rnd(W, D) for _ in range(NL)]
o_w = [rnd(D, D) for _ in range(NL)]
gu_w = [rnd(2 * F, D) for _ in range(NL)]
down_w = [rnd(D, F) for _ in range(NL)]
head_w = rnd(V, D); head_b = torch.randn(V, device=dev)
hn = rnd(B, D); att = rnd(B, Hq * dh); mid = rnd(B, F)
h = torch.randn(B, D, device=dev); qkv = torch.zeros(B, W, device=dev)
midout = torch.empty(B, F, device=dev, dtype=torch.bfloat16)
logits = torch.empty(B, V, device=dev)
ln_w = torch.zeros(D, device=dev)
hn_out = torch.empty(B, D, device=dev, dtype=torch.bfloat16)

def timeit(name, fn, bytes_per_launch, reps=3):
    for i in range(NL): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for i in range(NL): fn(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * NL)
    print(f"{name:34s} {us:8.2f} us   {bytes_per_launch / us / 1e3:8.1f} GB/s  ideal {bytes_per_launch / 6550.7e3:6.2f} us")

def graph_time(name, fn, bytes_per_launch, reps=5):
    """same, but the 18 launches captured in one CUDA graph (no host launch overhead)"""
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
    print(f"{name:34s} {us:8.2f} us   {bytes_per_launch / us / 1e3:8.1f} GB/s  ideal {bytes_per_launch / 6550.7e3:6.2f} us   [graph]")

# attention inputs
PAGE=64; kvlen=324; max_pages=7
k_pages=[rnd(B*max_pages, PAGE, dh) for _ in range(NL)]; v_pages=[rnd(B*max_pages, PAGE, dh) for _ in range(NL)]
table=torch.arange(B*max_pages, device=dev, dtype=torch.int32).view(B,max_pages).contiguous()
kvl=torch.full((B,), kvlen, device=dev, dtype=torch.int32); posd=torch.full((B,), kvlen, device=dev, dtype=torch.int32)
inv_freq=(1.0/(10000.0**(torch.arange(0,dh,2,dtype=torch.int64).float()/dh))).to(dev)
qkvf=torch.randn(B, W, device=dev)*0.5
attout=torch.empty(B, Hq*dh, device=dev, dtype=torch.bfloat16)
def attn(i):
    _lib.check(L.pg_attention_decode_fused(qkvf.data_ptr(), posd.data_ptr(), kvl.data_ptr(), inv_freq.data_ptr(), k_pages[i].data_ptr(), v_pages[i].data_ptr(), table.data_ptr(), attout.data_ptr(), B, Hq, Hkv, dh, PAGE, B*max_pages, max_pages, 1.0/16, _lib.stream()), "attn")
for pdl in (1,):
    L.pg_set_pdl(pdl)
    print(f"==== PDL={pdl}  B={B}")
    for sp in (2, 4, 7, 14):
        graph_time(f"qkv split{sp}", lambda i: _lib.gemm(hn, qkv_w[i], qkv, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sp), W * D * 2)
    for sp in (1, 2, 4, 9, 18):
        graph_time(f"o split{sp}", lambda i: _lib.gemm(att, o_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sp), D * D * 2)
    graph_time("gate-up geglu", lambda i: _lib.gemm(hn, gu_w[i], midout, mode=_lib.EPI_GEGLU, swap=1), 2 * F * D * 2)
    for sp in (4, 9, 18, 36):
        graph_time(f"down split{sp}", lambda i: _lib.gemm(mid, down_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sp), D * F * 2)
    graph_time("lm_head", lambda i: _lib.gemm(hn, head_w, logits, mode=_lib.EPI_F32, bias=head_b, swap=1), V * D * 2, reps=1)
    graph_time("rmsnorm", lambda i: _lib.rmsnorm(h, ln_w, hn_out), B * D * 6)
    graph_time("attention fused (kv 324)", attn, B * kvlen * dh * 2 * 2)
    def full_layer(i):
        _lib.rmsnorm(h, ln_w, hn_out, zero_buf=qkv)
        _lib.gemm(hn_out, qkv_w[i], qkv, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=7)
        attn(i)
        _lib.gemm(att, o_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=9)
        _lib.rmsnorm(h, ln_w, hn_out)
        _lib.gemm(hn_out, gu_w[i], midout, mode=_lib.EPI_GEGLU, swap=1)
        _lib.gemm(midout, down_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=18)
    graph_time("FULL layer", full_layer, (W * D + D * D + 3 * F * D) * 2 + B * kvlen * dh * 4)
    for sq, so, sd in ((7, 9, 36), (7, 9, 24), (14, 9, 18), (7, 18, 18), (5, 8, 18), (10, 9, 18)):
        def fl(i, sq=sq, so=so, sd=sd):
            _lib.rmsnorm(h, ln_w, hn_out, zero_buf=qkv)
            _lib.gemm(hn_out, qkv_w[i], qkv, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sq)
            attn(i)
            _lib.gemm(att, o_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=so)
            _lib.rmsnorm(h, ln_w, hn_out)
            _lib.gemm(hn_out, gu_w[i], midout, mode=_lib.EPI_GEGLU, swap=1)
            _lib.gemm(midout, down_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sd)
        graph_time(f"FULL layer splits qkv {sq} o {so} down {sd}", fl, (W * D + D * D + 3 * F * D) * 2 + B * kvlen * dh * 4)
    def layer(i):
        _lib.rmsnorm(h, ln_w, hn_out, zero_buf=qkv)
        _lib.gemm(hn_out, qkv_w[i], qkv, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=7)
        _lib.gemm(att, o_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=9)
        _lib.rmsnorm(h, ln_w, hn_out)
        _lib.gemm(hn_out, gu_w[i], midout, mode=_lib.EPI_GEGLU, swap=1)
        _lib.gemm(midout, down_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=9)
    graph_time("layer (no attention)", layer, (W * D + D * D + 3 * F * D) * 2)
    timeit("layer (no attention) eager", layer, (W * D + D * D + 3 * F * D) * 2)

    # ---- L2 prefetch experiments: pull part of the MLP weights into L2 while the latency-bound attention half runs ----
    def make_layer(pf_gu_mb, pf_down_mb, pf_next_attn):
        def f(i):
            gu = gu_w[i]; dn = down_w[i]
            _lib.rmsnorm(h, ln_w, hn_out, zero_buf=qkv, prefetch=gu if pf_gu_mb else None, prefetch_bytes=min(pf_gu_mb << 20, gu.numel() * 2))
            _lib.gemm(hn_out, qkv_w[i], qkv, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=7)
            attn(i)
            _lib.gemm(att, o_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=9)
            if pf_down_mb:
                _lib.rmsnorm(h, ln_w, hn_out, prefetch=dn, prefetch_bytes=min(pf_down_mb << 20, dn.numel() * 2))
            elif pf_next_attn:
                _lib.rmsnorm(h, ln_w, hn_out, prefetch=qkv_w[(i + 1) % NL])
            else:
                _lib.rmsnorm(h, ln_w, hn_out)
            _lib.gemm(hn_out, gu, midout, mode=_lib.EPI_GEGLU, swap=1)
            _lib.gemm(midout, dn, h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=18)
        return f
    nbytes = (W * D + D * D + 3 * F * D) * 2 + B * kvlen * dh * 4
    for gu_mb in (0, 16, 32, 48, 64, 96, 128):
        graph_time(f"FULL layer, L2-prefetch gu {gu_mb} MB", make_layer(gu_mb, 0, False), nbytes)
    graph_time("FULL layer, pf gu 64 MB + next qkv", make_layer(64, 0, True), nbytes)
    graph_time("FULL layer, pf gu 48 MB + down 32 MB", make_layer(48, 32, False), nbytes)
