import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
B, V = 64, 257216
cnt_ptr = 0
nxt = torch.zeros(B, device="cuda", dtype=torch.int32); cnt = torch.zeros(B, device="cuda", dtype=torch.int32)
step = torch.zeros(1, device="cuda", dtype=torch.int32)
def run(name, logits, reps=20):
    global cnt_ptr
    r0 = L.pg_debug_topp_retries()
    f = lambda: _lib.check(L.pg_sample_top_p(logits.data_ptr(), V, nxt.data_ptr(), cnt_ptr, B, V, 1.25, 0.9, 1234, step.data_ptr(), _lib.stream()), "topp")
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    import ctypes
    buf = (ctypes.c_longlong * 16)(); L.pg_debug_topp_trace(ctypes.addressof(buf))
    ph = [(buf[i + 1] - buf[i]) / 1.9e3 for i in range(5)]
    print("      phases us: P0 %.1f | PS %.1f | P1 %.1f | P2 %.1f | P3 %.1f" % tuple(ph))
    print(f"{name:28s} {e0.elapsed_time(e1) * 1e3 / reps:7.1f} us   kept median {cnt.float().median().item():9.0f}   retries/launch {(L.pg_debug_topp_retries() - r0) / (reps + 1):5.1f} of {B}")
import os
L.pg_debug_set_topp_bracket(int(os.environ.get("BRK", "0")))
cnt_ptr = cnt.data_ptr() if os.environ.get("CNT", "1") == "1" else 0
for std in (0.05, 0.3, 0.45, 1.0, 2.0, 4.0):
    run(f"randn * {std}", torch.randn(B, V, device="cuda") * std)
x = torch.randn(B, V, device="cuda") * 0.5; x[:, :50] += 12
run("peaked (50 hot tokens)", x)
