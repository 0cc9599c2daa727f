import math, os, sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (B, N) in ((64, 256), (32, 1024), (8, 4096)):
    H, dh = 16, 72; D = H * dh
    qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.7).bfloat16()
    out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    f = lambda: _lib.check(L.pg_attention_prefill(qkv.data_ptr(), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, out.data_ptr(), B, H, N, N, dh, 1,
        N * 3 * D, 3 * D, 0, dh, N * 3 * D, 3 * D, dh, N * D, D, 0, dh, dh ** -0.5, _lib.stream()), "a")
    ms = t(f); fl = 4.0 * B * H * N * N * dh
    print(f"siglip B={B} N={N}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s (algorithmic dh=72)")
    Hq, dh = 8, 256; S = N + 4
    q = (torch.randn(B * S, Hq * dh, device="cuda") * 0.5).bfloat16(); k = (torch.randn(B * S, dh, device="cuda") * 0.5).bfloat16(); v = (torch.randn(B * S, dh, device="cuda") * 0.5).bfloat16()
    o = torch.empty(B * S, Hq * dh, device="cuda", dtype=torch.bfloat16)
    g = lambda: _lib.check(L.pg_attention_prefill(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, 1, S * Hq, S, dh, Hq,
        S * Hq * dh, Hq * dh, dh, 0, S * dh, dh, 0, S * Hq * dh, Hq * dh, dh, 0, 1 / 16, _lib.stream()), "a")
    ms = t(g); fl = 4.0 * B * Hq * S * S * dh
    print(f"gemma  B={B} S={S}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
