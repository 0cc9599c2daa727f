"""Decode-layer chain (3B shapes, 64 sequences, 18 layers' weights in rotation) with L2 weight prefetch on FORKED graph
branches (pg_prefetch_l2): the gate||up weights are pulled into L2 while the attention block leaves the HBM idle, the down
weights while gate||up runs out of L2.   python profiles/tools/decode_l2_prefetch.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
B, D, F, Hq, Hkv, dh, NL = int(os.environ.get("MB_B", 64)), 2048, 16384, 8, 1, 256, 18
kvlen = int(os.environ.get("MB_KV", 324))
W = (Hq + 2 * Hkv) * dh
dev = "cuda"
PEAK = 6550.7e3
def rnd(*s): return (torch.randn(*s, device=dev) * 0.02).bfloat16()
qkv_w = [rnd(W, D) for _ in range(NL)]; o_w = [rnd(D, D) for _ in range(NL)]
gu_w = [rnd(2 * F, D) for _ in range(NL)]; down_w = [rnd(D, F) for _ in range(NL)]
h = torch.randn(B, D, device=dev); qkv = torch.zeros(B, W, device=dev)
midout = torch.empty(B, F, device=dev, dtype=torch.bfloat16)
ln_w = torch.zeros(D, device=dev); hn_out = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
PAGE = 64; max_pages = (kvlen + PAGE - 1) // PAGE + 1
k_pages = [rnd(B * max_pages, PAGE, dh) for _ in range(NL)]; v_pages = [rnd(B * max_pages, PAGE, dh) for _ in range(NL)]
table = torch.arange(B * max_pages, device=dev, dtype=torch.int32).view(B, max_pages).contiguous()
kvl = torch.full((B,), kvlen, device=dev, dtype=torch.int32); posd = kvl.clone()
inv_freq = (1.0 / (10000.0 ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))).to(dev)
attout = torch.empty(B, Hq * dh, device=dev, dtype=torch.bfloat16)
sms = _lib.num_sms(); SQ, SO, SD = max(1, sms // 20), max(1, sms // 16), max(1, 2 * sms // 16)
main = torch.cuda.Stream(); side = torch.cuda.Stream()
forked = [False]

def pf(t, mb, ctas=0, off_mb=0, el=0):
    o = int(off_mb * 1e6) & ~127
    n = min(t.numel() * 2 - o, int(mb * 1e6)) & ~15
    if n >= 16: _lib.check(L.pg_prefetch_l2(t.data_ptr() + o, n, ctas, el, _lib.stream()), "pf")

def attn(i):
    _lib.check(L.pg_attention_decode_fused(qkv.data_ptr(), posd.data_ptr(), kvl.data_ptr(), inv_freq.data_ptr(), k_pages[i].data_ptr(),
        v_pages[i].data_ptr(), table.data_ptr(), attout.data_ptr(), B, Hq, Hkv, dh, PAGE, B * max_pages, max_pages, 1.0 / 16, _lib.stream()), "attn")

def layer(i, plan=()):
    """plan: tuples (fork point, tensor name, MB, ctas, offset MB); fork points: 0 layer start, 1 after norm1, 2 after qkv,
    3 after attention, 4 after o_proj, 5 after norm2 (G start), 6 after G"""
    def forks(at):
        todo = [p for p in plan if p[0] == at]
        if not todo: return
        side.wait_stream(main)
        forked[0] = True
        with torch.cuda.stream(side):
            for _, name, mb, ctas, off, *rest in todo:
                pf({"g": gu_w, "d": down_w, "q": qkv_w, "o": o_w}[name][(i + (1 if name in "qo" and at >= 5 else 0)) % NL], mb, ctas, off, rest[0] if rest else 0)
    forks(0)
    _lib.rmsnorm(h, ln_w, hn_out)
    forks(1)
    _lib.gemm(hn_out, qkv_w[i], qkv, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=SQ)
    forks(2)
    attn(i)
    forks(3)
    _lib.gemm_fused(attout, o_w[i], h, mode=_lib.EPI_ATOMIC_F32, split_k=SO, zero_buf=qkv)
    forks(4)
    _lib.rmsnorm(h, ln_w, hn_out)
    forks(5)
    _lib.gemm(hn_out, gu_w[i], midout, mode=_lib.EPI_GEGLU, swap=1)
    forks(6)
    _lib.gemm(midout, down_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=SD)

layer_bytes = (W * D + D * D + 3 * F * D) * 2 + B * kvlen * dh * 4

def graph_time(name, fn, reps=5):
    with torch.cuda.stream(main):
        for i in range(NL): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    forked[0] = False
    with torch.cuda.graph(g, stream=main):
        for i in range(NL): fn(i)
        if forked[0]: main.wait_stream(side)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): g.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (reps * NL))
    print(f"{name:60s} {best:8.2f} us   {layer_bytes / best / 1e3:8.1f} GB/s  ideal {layer_bytes / PEAK:6.2f} us", flush=True)

print(f"==== B={B} kv={kvlen}")
graph_time("LAYER baseline", lambda i: layer(i))
for ctas in (32, 64, 148):
    for mb in (48, 64, 80):
        graph_time(f"G {mb} MB forked after qkv, {ctas} CTAs", lambda i: layer(i, ((2, "g", mb, ctas, 0),)))
for ctas in (64, 148):
    for mb in (32, 48, 64):
        graph_time(f"G {mb} MB forked after attention, {ctas} CTAs", lambda i: layer(i, ((3, "g", mb, ctas, 0),)))
graph_time("G 32 MB after qkv (32) + 32 MB after attention (148)", lambda i: layer(i, ((2, "g", 32, 32, 0), (3, "g", 32, 148, 32))))
sys.exit(0)
for el in (0, 1):
    graph_time(f"G 64 MB after qkv (64), evict_last={el}", lambda i: layer(i, ((2, "g", 64, 64, 0, el),)))
    for at, nm in ((0, "layer start"), (2, "after qkv"), (3, "after attention")):
        for ctas in (32, 64, 148):
            for mb in (34, 67):
                graph_time(f"D {mb} MB {nm}, {ctas} CTAs, evict_last={el}", lambda i: layer(i, ((at, "d", mb, ctas, 0, el),)))
    graph_time(f"G 48 MB after qkv (64) + D 34 MB after attention (64) el={el}", lambda i: layer(i, ((2, "g", 48, 64, 0, el), (3, "d", 34, 64, 0, el))))
    graph_time(f"D 67 MB after qkv (64) + G 24 MB after attention (64) el={el}", lambda i: layer(i, ((2, "d", 67, 64, 0, el), (3, "g", 24, 64, 0, el))))
