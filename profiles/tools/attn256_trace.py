"""Timeline of one CTA of the dh = 256 tcgen05 prefill attention (Gemma MQA, 8 x 4100 tokens): S issuer, P V issuer, softmax warp 2."""
import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
Hq, Hkv, dh, B, S = 8, 1, 256, 8, 4100
G = Hq // Hkv; W = (Hq + 2 * Hkv) * dh
qkv = (torch.randn(B * S, W, device="cuda") * 0.3).bfloat16()
att = torch.empty(B * S, Hq * dh, device="cuda", dtype=torch.bfloat16)
q, k, v = qkv, qkv[:, Hq * dh:], qkv[:, (Hq + Hkv) * dh:]
def launch():
    _lib.check(L.pg_attention_prefill(q.data_ptr(), k.data_ptr(), v.data_ptr(), att.data_ptr(), B, Hkv, S * G, S, dh, G,
                                      S * W, W, dh, G * dh, S * W, W, dh, S * Hq * dh, Hq * dh, dh, G * dh, dh ** -0.5, _lib.stream()), "a")
for _ in range(3): launch()
tr = torch.zeros(3 * 32 * 8, device="cuda", dtype=torch.int64)
L.pg_debug_set_attn_prefill_trace(tr.data_ptr())
launch(); torch.cuda.synchronize()
L.pg_debug_set_attn_prefill_trace(0)
t = tr.cpu().numpy().astype("float64").reshape(3, 32, 8)
t0 = t[0, 8, 0]
mma = ["S issuer loop top", "S(j) issued", "V landed", "ones planted", "P(j) ready", "PV(j) issued"]
sm = ["loop top", "S(j) ready", "S loaded", "max done", "exp+store done", "PV(j-1) done", "P(j) published"]
for j in range(10, 14):
    print(f"tile {j}")
    print("   issuers : " + " | ".join(f"{n} {int(t[0, j, e] - t0):6d}" for e, n in enumerate(mma)))
    print("   softmax : " + " | ".join(f"{n} {int(t[1, j, e] - t0):6d}" for e, n in enumerate(sm)))
print("softmax tile period (clk):", (t[1, 24, 0] - t[1, 8, 0]) / 16)
