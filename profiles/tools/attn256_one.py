"""The dh = 256 tcgen05 prefill attention (Gemma MQA: 8 query heads stacked as rows against 1 KV head) at the three prompt lengths:
64 x 260 / 32 x 1028 / 8 x 4100 tokens.  `python profiles/tools/attn256_one.py`"""
import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
Hq, Hkv, dh = 8, 1, 256
G = Hq // Hkv
W = (Hq + 2 * Hkv) * dh
for sw, (B, S) in [(sw, bs) for sw in (1, 2) for bs in ((64, 260), (32, 1028), (8, 4100))]:
    L.pg_debug_set_attn_prefill(sw << 4, -1)
    qkv = (torch.randn(B * S, W, device="cuda") * 0.3).bfloat16()
    att = torch.empty(B * S, Hq * dh, device="cuda", dtype=torch.bfloat16)
    q, k, v = qkv, qkv[:, Hq * dh:], qkv[:, (Hq + Hkv) * dh:]
    def launch():
        _lib.check(L.pg_attention_prefill(q.data_ptr(), k.data_ptr(), v.data_ptr(), att.data_ptr(), B, Hkv, S * G, S, dh, G,
                                          S * W, W, dh, G * dh, S * W, W, dh, S * Hq * dh, Hq * dh, dh, G * dh, dh ** -0.5, _lib.stream()), "a")
    for _ in range(3): launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): launch()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"SW={sw} dh=256 attention, {B} x {S} tokens x {Hq} query heads: {ms:.3f} ms/launch = {4.0 * B * Hq * S * S * dh / ms / 1e9:.0f} TFLOP/s", flush=True)
