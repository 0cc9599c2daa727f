"""B = 1 prefill (T = 256 / 260 tokens): N tile x split-K sweep of the token-major GEMMs, weights rotated over copies so
that they come from HBM as in the real layer loop."""
import sys, torch
sys.path.insert(0, '.')
from paligemma_multimodal_system_b200 import _lib
L = _lib.lib()
def rnd(*s): return (torch.randn(*s, device="cuda") * 0.05).bfloat16()
def t(fns, reps=4):
    for f in fns: f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for f in fns: f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(fns)) * 1e3
shapes = [("siglip qkv", 256, 3456, 1152, "bf16"), ("siglip out", 256, 1152, 1152, "f32r"), ("siglip fc1", 256, 4304, 1152, "gelu"),
          ("siglip fc2", 256, 1152, 4304, "f32r"), ("gemma qkv", 260, 2560, 2048, "bf16"), ("gemma o", 260, 2048, 2048, "f32r"),
          ("gemma gu", 260, 32768, 2048, "geglu"), ("gemma down", 260, 2048, 16384, "f32r")]
for name, T, F, K, kind in shapes:
    nc = max(2, min(12, int(400e6 / (2 * F * K))))
    x, ws = rnd(T, K), [rnd(F, K) for _ in range(nc)]
    bias = torch.randn(F, device="cuda")
    res = []
    for bn in (64, 128, 256):
        L.pg_debug_set_gemm_bn(bn)
        if kind == "f32r":
            out = torch.randn(T, F, device="cuda")
            for split in (1, 2, 3, 4, 6, 8, 12, 16):
                if split == 1:
                    fns = [lambda w=w: _lib.gemm(x, w, out, mode=_lib.EPI_F32, resid=out, swap=0) for w in ws]
                else:
                    if (K // 64) // split < 2: continue
                    fns = [lambda w=w: _lib.gemm(x, w, out, mode=_lib.EPI_ATOMIC_F32, swap=0, split_k=split) for w in ws]
                res.append((t(fns), bn, split))
        else:
            if kind == "geglu":
                out = torch.empty(T, F // 2, device="cuda", dtype=torch.bfloat16)
                fns = [lambda w=w: _lib.gemm(x, w, out, mode=_lib.EPI_GEGLU, swap=0) for w in ws]
            else:
                out = torch.empty(T, F, device="cuda", dtype=torch.bfloat16)
                fns = [lambda w=w: _lib.gemm(x, w, out, mode=_lib.EPI_BF16, bias=bias, act_gelu=(kind == "gelu"), swap=0) for w in ws]
            res.append((t(fns), bn, 1))
    L.pg_debug_set_gemm_bn(0)
    if kind == "f32r":
        out = torch.randn(T, F, device="cuda")
        cur = t([lambda w=w: _lib.gemm_residual(x, w, out) for w in ws])
    else:
        cur = t(fns)
    res.sort()
    print(f"{name:11s} T={T} F={F} K={K}: current policy {cur:6.1f} us | best " + ", ".join(f"{us:.1f}us(BN{bn},s{sp})" for us, bn, sp in res[:4])
          + f" | floor: weights {2 * F * K / 6.55e6:.1f} us, flops {2 * T * F * K / 1.387e9:.1f} us", flush=True)
