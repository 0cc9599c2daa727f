"""dh = 72 prefill attention, two query tiles per CTA: start offset of the second softmax group (after its first S tile)."""
import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
H, dh = 16, 72
D = H * dh
def run(N, reps=20):
    B = 32768 // N
    qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.7).bfloat16()
    out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    def launch():
        _lib.check(L.pg_attention_prefill(qkv.data_ptr(), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, out.data_ptr(), B, H, N, N, dh, 1,
                                          N * 3 * D, 3 * D, 0, dh, N * 3 * D, 3 * D, dh, N * D, D, 0, dh, dh ** -0.5, _lib.stream()), "a")
    for _ in range(3): launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): launch()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 4.0 * B * H * N * N * dh / ms / 1e9
for N in (4096, 1024):
    for stg in (0, 300, 500, 700, 900, 1100, 1400):
        L.pg_debug_set_attn_prefill(2, stg)
        ms, tf = run(N)
        print(f"N={N}: QT=2 stagger {stg:5d}: {ms:.3f} ms {tf:.0f} TF/s", flush=True)
L.pg_debug_set_attn_prefill(0, 0)
