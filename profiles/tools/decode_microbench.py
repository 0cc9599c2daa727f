"""Steady-state timing of the decode-layer kernel chain (3B shapes), 18 layers' weights in rotation (cold L2), every chain
captured in one CUDA graph (no host launch overhead).  A/B of the 7-launch chain (standalone RMSNorms, bf16 operands by TMA)
against the chains with the norms folded into the q/k/v and / or gate||up GEMMs.

    python profiles/tools/decode_microbench.py            # MB_B=64 MB_KV=324 by default
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from paligemma_multimodal_system_b200 import _lib  # noqa: E402

L = _lib.lib()
B, D, F, Hq, Hkv, dh, V, NL = int(os.environ.get("MB_B", 64)), 2048, 16384, 8, 1, 256, 257216, 18
kvlen = int(os.environ.get("MB_KV", 324))
W = (Hq + 2 * Hkv) * dh
dev = "cuda"
PEAK = 6550.7e3  # bytes / us


def rnd(*s):
    return (torch.randn(*s, device=dev) * 0.02).bfloat16()


qkv_w = [rnd(W, D) for _ in range(NL)]
o_w = [rnd(D, D) for _ in range(NL)]
gu_w = [rnd(2 * F, D) for _ in range(NL)]
down_w = [rnd(D, F) for _ in range(NL)]
head_w = rnd(V, D)
head_b = torch.randn(V, device=dev)
hn = rnd(B, D)
att = rnd(B, Hq * dh)
mid = rnd(B, F)
h = torch.randn(B, D, device=dev)
qkv = torch.zeros(B, W, device=dev)
midout = torch.empty(B, F, device=dev, dtype=torch.bfloat16)
logits = torch.empty(B, V, device=dev)
ln_w = torch.zeros(D, device=dev)
hn_out = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
PAGE = 64
max_pages = (kvlen + PAGE - 1) // PAGE + 1
k_pages = [rnd(B * max_pages, PAGE, dh) for _ in range(NL)]
v_pages = [rnd(B * max_pages, PAGE, dh) for _ in range(NL)]
table = torch.arange(B * max_pages, device=dev, dtype=torch.int32).view(B, max_pages).contiguous()
kvl = torch.full((B,), kvlen, device=dev, dtype=torch.int32)
posd = torch.full((B,), kvlen, device=dev, dtype=torch.int32)
inv_freq = (1.0 / (10000.0 ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))).to(dev)
attout = torch.empty(B, Hq * dh, device=dev, dtype=torch.bfloat16)
sms = _lib.num_sms()
SQ, SO, SD = max(1, sms // 20), max(1, sms // 16), max(1, 2 * sms // 16)


def graph_time(name, fn, bytes_per_launch, reps=5):
    for i in range(NL):
        fn(i)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(NL):
                fn(i)
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * NL)
    print(f"{name:52s} {us:8.2f} us   {bytes_per_launch / us / 1e3:8.1f} GB/s  ideal {bytes_per_launch / PEAK:6.2f} us", flush=True)
    return us


def attn(i):
    _lib.check(L.pg_attention_decode_fused(qkv.data_ptr(), posd.data_ptr(), kvl.data_ptr(), inv_freq.data_ptr(), k_pages[i].data_ptr(),
                                           v_pages[i].data_ptr(), table.data_ptr(), attout.data_ptr(), B, Hq, Hkv, dh, PAGE, B * max_pages,
                                           max_pages, 1.0 / 16, _lib.stream()), "attn")


def layer(i, sq=SQ, so=SO, sd=SD):
    _lib.rmsnorm(h, ln_w, hn_out)
    _lib.gemm(hn_out, qkv_w[i], qkv, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sq)
    attn(i)
    _lib.gemm_fused(attout, o_w[i], h, mode=_lib.EPI_ATOMIC_F32, split_k=so, zero_buf=qkv)
    _lib.rmsnorm(h, ln_w, hn_out)
    _lib.gemm(hn_out, gu_w[i], midout, mode=_lib.EPI_GEGLU, swap=1)
    _lib.gemm(midout, down_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sd)


layer_bytes = (W * D + D * D + 3 * F * D) * 2 + B * kvlen * dh * 4
print(f"==== B={B} kv={kvlen} splits qkv {SQ} o {SO} down {SD}")
graph_time("qkv split-K", lambda i: _lib.gemm(hn, qkv_w[i], qkv, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=SQ), W * D * 2)
graph_time("o   split-K (+ zero-fill of qkv)", lambda i: _lib.gemm_fused(att, o_w[i], h, mode=_lib.EPI_ATOMIC_F32, split_k=SO, zero_buf=qkv), D * D * 2)
graph_time("gate-up geglu", lambda i: _lib.gemm(hn, gu_w[i], midout, mode=_lib.EPI_GEGLU, swap=1), 2 * F * D * 2)
for sd in (SD, 2 * SD, 37):
    graph_time(f"down split-K {sd}", lambda i, sd=sd: _lib.gemm(mid, down_w[i], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sd), D * F * 2)
graph_time("rmsnorm", lambda i: _lib.rmsnorm(h, ln_w, hn_out), B * D * 6)
graph_time("attention", lambda i: attn(i), B * kvlen * dh * 4)
graph_time("LAYER (7 launches)", layer, layer_bytes)
graph_time("LAYER, down split 37", lambda i: layer(i, sd=37), layer_bytes)

# ---- lm_head and the samplers (per decode step, not per layer: one launch per graph node, 18 nodes) ----
stats = torch.empty(4 * ((V + 127) // 128), B, 2, device=dev)
nxt = torch.empty(B, device=dev, dtype=torch.int32)
step = torch.zeros(1, device=dev, dtype=torch.int32)
graph_time("lm_head (fp32 logits + bias)", lambda i: _lib.gemm(hn, head_w, logits, mode=_lib.EPI_F32, bias=head_b, swap=1), V * D * 2, reps=1)
graph_time("lm_head + segment statistics", lambda i: _lib.gemm_fused(hn, head_w, logits, mode=_lib.EPI_F32, bias=head_b, stats=stats, inv_temperature=1.25), V * D * 2, reps=1)
logits.normal_(0, 2.0)
_lib.gemm_fused(hn, head_w, logits, mode=_lib.EPI_F32, bias=head_b, stats=stats, inv_temperature=1.25)
graph_time("top-p, 3 passes over the row (pg_sample_top_p)", lambda i: _lib.check(L.pg_sample_top_p(logits.data_ptr(), V, nxt.data_ptr(), 0, B, V, 1.25, 0.9, 1, step.data_ptr(), _lib.stream()), "s"), B * V * 4, reps=3)
graph_time("top-p from statistics (pg_sample_top_p_stats)", lambda i: _lib.check(L.pg_sample_top_p_stats(logits.data_ptr(), V, stats.data_ptr(), stats.shape[1], nxt.data_ptr(), B, V, 1.25, 0.9, 1, 0, step.data_ptr(), _lib.stream()), "s"), B * V * 4, reps=3)
for r in (1, 2):
    L.pg_debug_set_sampler_cluster(r)
    graph_time(f"top-p from statistics, {r} CTA(s) per row", lambda i: _lib.check(L.pg_sample_top_p_stats(logits.data_ptr(), V, stats.data_ptr(), stats.shape[1], nxt.data_ptr(), B, V, 1.25, 0.9, 1, 0, step.data_ptr(), _lib.stream()), "s"), B * V * 4, reps=3)
L.pg_debug_set_sampler_cluster(0)
graph_time("argmax over the row (pg_argmax)", lambda i: _lib.check(L.pg_argmax(logits.data_ptr(), V, nxt.data_ptr(), B, V, _lib.stream()), "a"), B * V * 4, reps=3)
graph_time("argmax from statistics (pg_argmax_stats)", lambda i: _lib.check(L.pg_argmax_stats(logits.data_ptr(), V, stats.data_ptr(), stats.shape[1], nxt.data_ptr(), B, V, _lib.stream()), "a"), B * V * 4, reps=3)


def head_and_sample(i):
    _lib.gemm_fused(hn, head_w, logits, mode=_lib.EPI_F32, bias=head_b, stats=stats, inv_temperature=1.25)
    _lib.check(L.pg_sample_top_p_stats(logits.data_ptr(), V, stats.data_ptr(), stats.shape[1], nxt.data_ptr(), B, V, 1.25, 0.9, 1, 0, step.data_ptr(), _lib.stream()), "s")


def head_and_sample_old(i):
    _lib.gemm(hn, head_w, logits, mode=_lib.EPI_F32, bias=head_b, swap=1)
    _lib.check(L.pg_sample_top_p(logits.data_ptr(), V, nxt.data_ptr(), 0, B, V, 1.25, 0.9, 1, step.data_ptr(), _lib.stream()), "s")


graph_time("lm_head -> top-p, round-1 kernels", head_and_sample_old, V * D * 2, reps=1)
graph_time("lm_head + statistics -> top-p from statistics", head_and_sample, V * D * 2, reps=1)
