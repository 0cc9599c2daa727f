"""Medium token counts (prefill groups of 8-32 images): wave quantisation of the 128x256 tiles on 148 SMs.  N tile sweep."""
import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
def rnd(*s): return (torch.randn(*s, device="cuda") * 0.05).bfloat16()
def t(fns, reps=3):
    for f in fns: f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for f in fns: f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(fns)) * 1e3
for nimg in (8, 16, 32):
    Ts, Tg = nimg * 256, nimg * 264
    shapes = [("siglip qkv", Ts, 3456, 1152, "bf16"), ("siglip out", Ts, 1152, 1152, "f32r"), ("siglip fc1", Ts, 4304, 1152, "gelu"),
              ("siglip fc2", Ts, 1152, 4304, "f32r"), ("gemma qkv", Tg, 2560, 2048, "bf16"), ("gemma o", Tg, 2048, 2048, "f32r"),
              ("gemma gu", Tg, 32768, 2048, "geglu"), ("gemma down", Tg, 2048, 16384, "f32r")]
    for name, T, F, K, kind in shapes:
        nc = max(2, min(8, int(300e6 / (2 * F * K))))
        x, ws = rnd(T, K), [rnd(F, K) for _ in range(nc)]
        bias = torch.randn(F, device="cuda")
        res = []
        for bn in (0, 128, 256):
            L.pg_debug_set_gemm_bn(bn)
            if kind == "f32r":
                out = torch.randn(T, F, device="cuda")
                fns = [lambda w=w: _lib.gemm(x, w, out, mode=_lib.EPI_F32, resid=out, swap=0) for w in ws]
            elif kind == "geglu":
                out = torch.empty(T, F // 2, device="cuda", dtype=torch.bfloat16)
                fns = [lambda w=w: _lib.gemm(x, w, out, mode=_lib.EPI_GEGLU, swap=0) for w in ws]
            else:
                out = torch.empty(T, F, device="cuda", dtype=torch.bfloat16)
                fns = [lambda w=w: _lib.gemm(x, w, out, mode=_lib.EPI_BF16, bias=bias, act_gelu=(kind == "gelu"), swap=0) for w in ws]
            res.append((bn, t(fns)))
        L.pg_debug_set_gemm_bn(0)
        mt = (T + 127) // 128
        print(f"{nimg:2d} img {name:11s} T={T} tiles256={mt * ((F + 255) // 256):5d} ({mt * ((F + 255) // 256) / 148:.2f} waves): " +
              "  ".join(f"BN{bn if bn else 'auto'} {us:7.1f}us" for bn, us in res), flush=True)
