"""CTA-pair (cta_group::2) prefill GEMM vs the one-CTA kernel on the PaliGemma prefill shapes: bitwise comparison + timing."""
import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
def t(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def rnd(*s): return (torch.randn(*s, device="cuda") * 0.05).bfloat16()
T1, T2 = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, int(sys.argv[2]) if len(sys.argv) > 2 else 16640
shapes = [("siglip qkv   bf16+bias", T1, 3456, 1152, "bf16b"), ("siglip out   f32+bias+resid", T1, 1152, 1152, "f32r"),
          ("siglip fc1   bf16+bias+gelu", T1, 4304, 1152, "gelu"), ("siglip fc2   f32+bias+resid", T1, 1152, 4304, "f32r"),
          ("projector    f32", T1, 2048, 1152, "f32"),
          ("gemma qkv    bf16", T2, 2560, 2048, "bf16"), ("gemma o      f32+resid", T2, 2048, 2048, "f32r"),
          ("gemma gu     geglu", T2, 32768, 2048, "geglu"), ("gemma down   f32+resid", T2, 2048, 16384, "f32r"),
          ("ragged       bf16+bias", 16384 + 77, 3456 + 8, 1152, "bf16b")]
tot = [0.0, 0.0, 0.0, 0.0]
for name, T, F, K, kind in shapes:
    x, w = rnd(T, K), rnd(F, K)
    bias = torch.randn(F, device="cuda")
    res0 = torch.randn(T, F, device="cuda") if kind == "f32r" else None
    def make(out):
        if kind in ("bf16b", "bf16", "gelu"):
            return lambda: _lib.gemm(x, w, out, mode=_lib.EPI_BF16, bias=None if kind == "bf16" else bias, act_gelu=(kind == "gelu"), swap=0)
        if kind == "geglu":
            return lambda: _lib.gemm(x, w, out, mode=_lib.EPI_GEGLU, swap=0)
        if kind == "f32r":
            return lambda: _lib.gemm_residual(x, w, out, bias=bias if "siglip" in name else None)
        return lambda: _lib.gemm(x, w, out, mode=_lib.EPI_F32, swap=0)
    outs, mss = [], []
    for pair in (2, 3, 0, 1):  # bit 0: CTA pair, bit 1: banded raster OFF
        L.pg_debug_set_gemm_pair(pair, 0)
        if kind in ("bf16b", "bf16", "gelu"): out = torch.zeros(T, F, device="cuda", dtype=torch.bfloat16)
        elif kind == "geglu": out = torch.zeros(T, F // 2, device="cuda", dtype=torch.bfloat16)
        elif kind == "f32r": out = res0.clone()
        else: out = torch.zeros(T, F, device="cuda")
        fn = make(out)
        fn(); torch.cuda.synchronize()
        outs.append(out.clone())
        mss.append(t(fn))
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    fl = 2.0 * T * F * K
    for i in range(4): tot[i] += mss[i]
    print(f"{name:30s} T={T} F={F:6d} K={K:6d}: one-CTA {mss[0] * 1e3:8.1f} us {fl / mss[0] / 1e9:5.0f} TF/s | pair {mss[1] * 1e3:8.1f} us {fl / mss[1] / 1e9:5.0f} | one-CTA banded {mss[2] * 1e3:8.1f} us {fl / mss[2] / 1e9:5.0f} | pair banded {mss[3] * 1e3:8.1f} us {fl / mss[3] / 1e9:5.0f} | bitwise equal {same}", flush=True)
print("sum: " + ", ".join(f"{n} {t * 1e3:.1f} us" for n, t in zip(("one-CTA", "pair", "one-CTA banded", "pair banded"), tot)))
L.pg_debug_set_gemm_pair(1, 0)
