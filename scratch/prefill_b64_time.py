"""Device time of the 64-image prefill (vision | merge + decoder), median of 7, for A/B runs of GEMM policies."""
import sys, torch
sys.path.insert(0, '.')
from bench import PROMPT_LEN, build_gpu_model
from paligemma_multimodal_system_b200.modeling_gemma import KVCache
from paligemma_multimodal_system_b200.random_init import make_inputs, paligemma_3b_config
cfg = paligemma_3b_config(224)
model, _ = build_gpu_model(cfg)
lm = model.language_model
B = 64
inp = {k: v.cuda() for k, v in make_inputs(cfg, batch=B, prompt_len=PROMPT_LEN, seed=100).items()}
S = inp["input_ids"].shape[1]
kv = KVCache(); kv.allocate(B, 18, 1, 256, S + 64)
res = []
for it in range(10):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    e[0].record()
    img = model.image_features(inp["pixel_values"])
    e[1].record()
    h, pos = model._merge(inp["input_ids"], inp["attention_mask"], img)
    lm.prefill(h, pos, B, S, kv, last_only=True)
    e[2].record()
    torch.cuda.synchronize()
    if it >= 3:
        res.append((e[0].elapsed_time(e[2]), e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
res.sort()
t, v, g = res[len(res) // 2]
print(f"prefill 64 images: {t:.2f} ms ({t / B:.3f} ms/image)  vision {v:.2f} ms  merge+gemma {g:.2f} ms")
