"""Prefill time per image at the three BASELINE geometries (3B widths, HBM-resident inputs, CUDA events), with the
CTA-pair (cta_group::2) GEMM off / on.   python profiles/tools/prefill_bench.py [224 448 896]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import PROMPT_LEN, algorithmic_flops_per_image, build_gpu_model, peaks  # noqa: E402
from paligemma_multimodal_system_b200.modeling_gemma import KVCache  # noqa: E402
from paligemma_multimodal_system_b200.random_init import make_inputs, paligemma_3b_config  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [224, 448, 896]
tf_peak = peaks()[2]
for size in sizes:
    B = {224: 64, 448: 32, 896: 8}[size]
    cfg = paligemma_3b_config(size)
    model, _ = build_gpu_model(cfg)
    inp = {k: v.cuda() for k, v in make_inputs(cfg, batch=B, prompt_len=PROMPT_LEN, seed=100).items()}
    S = inp["input_ids"].shape[1]
    from paligemma_multimodal_system_b200 import _lib
    for fused in (False, True, False, True):  # (A/B switch reused: False = one-CTA GEMM only, True = CTA-pair GEMM)
        _lib.lib().pg_debug_set_gemm_pair(1 if fused else 3, 0)  # bit 1 set = banded raster OFF
        kv = KVCache()
        for _ in range(2):
            model.forward(inp["input_ids"], inp["pixel_values"], inp["attention_mask"], kv_cache=None, last_only=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 8
        e0.record()
        for _ in range(reps):
            kv = KVCache(reserve_tokens=8)
            model.forward(inp["input_ids"], inp["pixel_values"], inp["attention_mask"], kv_cache=kv, last_only=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps / B
        fl = algorithmic_flops_per_image(cfg, S)
        print(f"{size} px x {B}: banded_raster={fused}: {ms:.3f} ms/image = {fl / ms / 1e9:.0f} TFLOP/s = {fl / ms / 1e9 / tf_peak:.3f} of sustained peak", flush=True)
    del model
    torch.cuda.empty_cache()
