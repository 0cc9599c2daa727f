import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
def t(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def rnd(*s): return (torch.randn(*s, device="cuda") * 0.05).bfloat16()
T1, T2 = 16384, 16640
shapes = [("siglip qkv   bf16+bias", T1, 3456, 1152, "bf16b"), ("siglip out   f32+bias+resid", T1, 1152, 1152, "f32r"),
          ("siglip fc1   bf16+bias+gelu", T1, 4304, 1152, "gelu"), ("siglip fc2   f32+bias+resid", T1, 1152, 4304, "f32r"),
          ("projector    f32", T1, 2048, 1152, "f32"),
          ("gemma qkv    bf16", T2, 2560, 2048, "bf16"), ("gemma o      f32+resid", T2, 2048, 2048, "f32r"),
          ("gemma gu     geglu", T2, 32768, 2048, "geglu"), ("gemma down   f32+resid", T2, 2048, 16384, "f32r")]
tot = 0
for name, T, F, K, kind in shapes:
    x, w = rnd(T, K), rnd(F, K)
    bias = torch.randn(F, device="cuda")
    if kind in ("bf16b", "bf16", "gelu"):
        out = torch.empty(T, F, device="cuda", dtype=torch.bfloat16)
        fn = lambda: _lib.gemm(x, w, out, mode=_lib.EPI_BF16, bias=None if kind == "bf16" else bias, act_gelu=(kind == "gelu"), swap=0)
    elif kind == "geglu":
        out = torch.empty(T, F // 2, device="cuda", dtype=torch.bfloat16)
        fn = lambda: _lib.gemm(x, w, out, mode=_lib.EPI_GEGLU, swap=0)
    else:
        out = torch.randn(T, F, device="cuda")
        if kind == "f32r":
            fn = lambda: _lib.gemm_residual(x, w, out, bias=bias if "siglip" in name else None)
        else:
            fn = lambda: _lib.gemm(x, w, out, mode=_lib.EPI_F32, swap=0)
    ms = t(fn)
    fl = 2.0 * T * F * K
    y = torch.empty(T, F, device="cuda", dtype=torch.bfloat16)
    ms_ref = t(lambda: torch.matmul(x, w.t(), out=y))
    print(f"{name:30s} T={T} F={F:6d} K={K:6d}: {ms * 1e3:8.1f} us {fl / ms / 1e9:7.0f} TFLOP/s   | torch.matmul {ms_ref * 1e3:8.1f} us {fl / ms_ref / 1e9:7.0f} TFLOP/s")
