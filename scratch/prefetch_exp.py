"""Side-stream L2 warm-up experiment: while the latency-bound attention half of layer i runs on the main stream, a
second stream pulls X MB of layer i's gate||up (and optionally down) weights into L2."""
import os, sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
B, D, F, Hq, Hkv, dh, NL = 64, 2048, 16384, 8, 1, 256, 18
W = (Hq + 2 * Hkv) * dh
dev = "cuda"
def rnd(*s): return (torch.randn(*s, device=dev) * 0.02).bfloat16()
qkv_w = [rnd(W, D) for _ in range(NL)]; o_w = [rnd(D, D) for _ in range(NL)]
gu_w = [rnd(2 * F, D) for _ in range(NL)]; down_w = [rnd(D, F) for _ in range(NL)]
hn = rnd(B, D); att = rnd(B, Hq * dh)
h = torch.randn(B, D, device=dev); qkv = torch.zeros(B, W, device=dev)
midout = torch.empty(B, F, device=dev, dtype=torch.bfloat16)
ln_w = torch.zeros(D, device=dev); hn_out = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
PAGE = 64; kvlen = 324; max_pages = 7
k_pages = [rnd(B * max_pages, PAGE, dh) for _ in range(NL)]; v_pages = [rnd(B * max_pages, PAGE, dh) for _ in range(NL)]
table = torch.arange(B * max_pages, device=dev, dtype=torch.int32).view(B, max_pages).contiguous()
kvl = torch.full((B,), kvlen, device=dev, dtype=torch.int32); posd = torch.full((B,), kvlen, device=dev, dtype=torch.int32)
inv_freq = (1.0 / (10000.0 ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))).to(dev)
qkvf = torch.randn(B, W, device=dev) * 0.5
attout = torch.empty(B, Hq * dh, device=dev, dtype=torch.bfloat16)
def attn(i):
    _lib.check(L.pg_attention_decode_fused(qkvf.data_ptr(), posd.data_ptr(), kvl.data_ptr(), inv_freq.data_ptr(), k_pages[i].data_ptr(), v_pages[i].data_ptr(), table.data_ptr(), attout.data_ptr(), B, Hq, Hkv, dh, PAGE, B * max_pages, max_pages, 1.0 / 16, _lib.stream()), "attn")
nbytes = (W * D + D * D + 3 * F * D) * 2 + B * kvlen * dh * 4

def run(name, gu_mb, down_mb, mode, ctas, where="start", reps=5):
    side = torch.cuda.Stream()
    def layer(i):
        main = torch.cuda.current_stream()
        def pf():
            if gu_mb or down_mb:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    if gu_mb:
                        _lib.check(L.pg_prefetch_l2(gu_w[i].data_ptr(), min(gu_mb << 20, gu_w[i].numel() * 2), mode, ctas, side.cuda_stream), "pf")
                    if down_mb:
                        _lib.check(L.pg_prefetch_l2(down_w[i].data_ptr(), min(down_mb << 20, down_w[i].numel() * 2), mode, ctas, side.cuda_stream), "pf")
        if where == "start": pf()
        _lib.rmsnorm(h, ln_w, hn_out, zero_buf=qkv)
        if where == "after_ln": pf()
        _lib.gemm(hn_out, qkv_w[i], qkv, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=7)
        attn(i)
        _lib.gemm(att, o_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=9)
        _lib.rmsnorm(h, ln_w, hn_out)
        _lib.gemm(hn_out, gu_w[i], midout, mode=_lib.EPI_GEGLU, swap=1)
        _lib.gemm(midout, down_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=18)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(NL): layer(i)
        if gu_mb or down_mb: s.wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(NL): layer(i)
            if gu_mb or down_mb: s.wait_stream(side)
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * NL)
    print(f"{name:50s} {us:8.2f} us/layer  {nbytes / us / 1e3:8.1f} GB/s", flush=True)

L.pg_set_pdl(1)
run("baseline (no prefetch)", 0, 0, 0, 148)
for mode in (0, 1, 2):
    for gu_mb in (32, 64, 96):
        run(f"mode {mode} gu {gu_mb} MB, 148 ctas", gu_mb, 0, mode, 148)
run("mode 0 gu 64 MB, 296 ctas", 64, 0, 0, 296)
run("mode 0 gu 64 MB, 74 ctas", 64, 0, 0, 74)
run("mode 0 gu 64 MB, 32 ctas", 64, 0, 0, 32)
run("mode 0 gu 64 + down 32 MB", 64, 32, 0, 148)
run("mode 0 gu 96 MB after_ln", 96, 0, 0, 148, where="after_ln")
run("mode 2 gu 64 MB, 32 ctas", 64, 0, 2, 32)
