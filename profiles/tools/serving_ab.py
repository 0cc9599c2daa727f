"""Continuous batching (bench.measure_serving) with the q/k/v-epilogue RoPE path forced off / on."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import build_gpu_model, measure_serving  # noqa: E402
from paligemma_multimodal_system_b200.random_init import paligemma_3b_config  # noqa: E402

cfg = paligemma_3b_config(224)
model, _ = build_gpu_model(cfg)
for fused in (None, True, None, True, None, False):
    model.language_model.fused_qkv_rope = fused
    r = measure_serving(model, cfg, static=False)
    print(f"fused_qkv_rope={fused}: {r['useful_tokens_per_s']:.0f} useful tok/s, host_ms {r['host_ms']}", flush=True)
