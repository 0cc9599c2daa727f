"""The dh = 72 tcgen05 prefill attention (SigLIP) at the three image sizes, 32768 tokens per launch: 128 x 256 / 32 x 1024 /
8 x 4096 tokens, 16 heads.  `python profiles/tools/attn72_one.py [N]` runs one shape three times (for ncu); without an argument
it times all three with CUDA events."""
import sys

import torch

sys.path.insert(0, '.')  # run from the repo root
from paligemma_multimodal_system_b200 import _lib  # noqa: E402

L = _lib.lib()
H, dh = 16, 72
D = H * dh


def run(N, reps):
    B = 32768 // N
    qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.7).bfloat16()
    out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)

    def launch():
        _lib.check(L.pg_attention_prefill(qkv.data_ptr(), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, out.data_ptr(), B, H, N, N, dh, 1,
                                          N * 3 * D, 3 * D, 0, dh, N * 3 * D, 3 * D, dh, N * D, D, 0, dh, dh ** -0.5, _lib.stream()), "a")

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    if reps == 0:
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 4.0 * B * H * N * N * dh
    print(f"dh=72 attention, {B} x {N} tokens x {H} heads: {ms:.3f} ms/launch = {flops / ms / 1e9:.0f} TFLOP/s")


if len(sys.argv) > 1:
    run(int(sys.argv[1]), 0)
    print("ok")
else:
    for n in (256, 1024, 4096):
        run(n, 20)
