"""Timeline of one CTA of the dh = 72 tcgen05 prefill attention (8 x 4096 tokens x 16 heads): clock64 stamps of the MMA issuer and
of two softmax warps (column halves of the same rows) over consecutive key tiles."""
import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
H, dh, N = 16, 72, 4096
D = H * dh; B = 32768 // N
qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.7).bfloat16()
out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
def launch():
    _lib.check(L.pg_attention_prefill(qkv.data_ptr(), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, out.data_ptr(), B, H, N, N, dh, 1,
                                      N * 3 * D, 3 * D, 0, dh, N * 3 * D, 3 * D, dh, N * D, D, 0, dh, dh ** -0.5, _lib.stream()), "a")
QT = int(sys.argv[1]) if len(sys.argv) > 1 else 0
L.pg_debug_set_attn_prefill(QT, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
for _ in range(3): launch()
tr = torch.zeros(3 * 32 * 8, device="cuda", dtype=torch.int64)
L.pg_debug_set_attn_prefill_trace(tr.data_ptr())
launch(); torch.cuda.synchronize()
L.pg_debug_set_attn_prefill_trace(0)
t = tr.cpu().numpy().astype("float64").reshape(3, 32, 8)
t0 = t[0, 8, 0]
mma = ["loop top", "S(j+1) issued", "V landed", "ones planted", "P(j) ready", "PV(j) issued"]
sm = ["loop top", "S(j) ready", "S loaded", "max done", "exp+store done", "PV(j-1) done", "P(j) published"]
for j in range(8, 14):
    print(f"tile {j}")
    print("   MMA     : " + " | ".join(f"{n} {int(t[0, j, e] - t0):6d}" for e, n in enumerate(mma)))
    for role in (1, 2):
        print(f"   softmax{role}: " + " | ".join(f"{n} {int(t[role, j, e] - t0):6d}" for e, n in enumerate(sm)))
print("softmax tile period (clk):", (t[1, 24, 0] - t[1, 8, 0]) / 16)
