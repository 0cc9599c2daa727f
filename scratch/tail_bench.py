"""Times the non-layer part of a decode step (final norm, lm_head, sampler, advance, embed) in a CUDA graph."""
import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
B, D, V = 64, 2048, 257216
dev = "cuda"
def rnd(*s): return (torch.randn(*s, device=dev) * 0.02).bfloat16()
head_w = rnd(V, D); head_b = torch.randn(V, device=dev)
h = torch.randn(B, D, device=dev); hn = torch.empty(B, D, device=dev, dtype=torch.bfloat16); ln_w = torch.zeros(D, device=dev)
logits = torch.randn(B, V, device=dev)
nxt = torch.zeros(B, device=dev, dtype=torch.int32); cur = torch.zeros(B, device=dev, dtype=torch.int32)
hist = torch.zeros(256, B, device=dev, dtype=torch.int32); step = torch.zeros(1, device=dev, dtype=torch.int32)
counters = torch.zeros(3, B, device=dev, dtype=torch.int32)
def graph_time(name, fn, reps=20):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step.zero_()
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) * 1e3 / reps:8.2f} us", flush=True)
def norm(): _lib.rmsnorm(h, ln_w, hn)
def head(): _lib.gemm(hn, head_w, logits, mode=_lib.EPI_F32, bias=head_b, swap=1)
def argmax(): _lib.check(L.pg_argmax(logits.data_ptr(), V, nxt.data_ptr(), B, V, _lib.stream()), "argmax")
def topp(): _lib.check(L.pg_sample_top_p(logits.data_ptr(), V, nxt.data_ptr(), 0, B, V, 1.25, 0.9, 1234, step.data_ptr(), _lib.stream()), "topp")
def adv(): _lib.check(L.pg_advance_decode(nxt.data_ptr(), hist.data_ptr(), cur.data_ptr(), counters.data_ptr(), 3, step.data_ptr(), B, _lib.stream()), "adv")
def embed(): _lib.check(L.pg_embed_tokens(cur.data_ptr(), head_w.data_ptr(), 0, h.data_ptr(), B, D, 0, 45.0, 1.0, -1, -1, _lib.stream()), "embed")
graph_time("empty-ish graph (advance only)", adv)
graph_time("norm", norm)
graph_time("head", head)
graph_time("argmax", argmax)
graph_time("top-p", topp)
graph_time("embed", embed)
graph_time("norm+head", lambda: (norm(), head()))
graph_time("norm+head+argmax+adv+embed", lambda: (norm(), head(), argmax(), adv(), embed()))
graph_time("norm+head+topp+adv+embed", lambda: (norm(), head(), topp(), adv(), embed()))
# realistic logits from the head (diffuse)
head(); torch.cuda.synchronize()
graph_time("top-p on head logits", topp)
graph_time("argmax on head logits", argmax)
