"""cfg-2: PaliGemma-3B-224, batch 1, greedy 32 tokens (latency path)."""
import sys, torch
sys.path.insert(0, '.')
from bench import PROMPT_LEN, build_gpu_model
from paligemma_multimodal_system_b200.random_init import make_inputs, paligemma_3b_config
cfg = paligemma_3b_config(224)
model, _ = build_gpu_model(cfg)
for B in (1, 8):
    inp = {k: v.cuda() for k, v in make_inputs(cfg, batch=B, prompt_len=PROMPT_LEN, seed=100).items()}
    for _ in range(3):
        model.generate(inp["input_ids"], inp["pixel_values"], inp["attention_mask"], 32, do_sample=False)
    res = []
    for _ in range(5):
        tm = {}
        model.generate(inp["input_ids"], inp["pixel_values"], inp["attention_mask"], 32, do_sample=False, timings=tm)
        res.append(tm)
    pre = sorted(r["prefill_ms"] for r in res)[2]; dec = sorted(r["decode_ms"] for r in res)[2]
    print(f"B={B}: prefill {pre:.2f} ms ({pre / B:.2f} ms/image), decode {dec / 31 * 1e3:.0f} us/token-step -> {B * 31 / dec * 1e3:.0f} tok/s; 32 tokens end to end {pre + dec:.1f} ms")
