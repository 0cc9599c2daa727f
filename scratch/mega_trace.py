import os, sys, torch
sys.path.insert(0, '.')
from bench import PROMPT_LEN, build_gpu_model
from paligemma_multimodal_system_b200.random_init import make_inputs, paligemma_3b_config
cfg = paligemma_3b_config(224)
model, _ = build_gpu_model(cfg)
B = int(os.environ.get("MB_B", 64))
inp = {k: v.cuda() for k, v in make_inputs(cfg, batch=B, prompt_len=PROMPT_LEN, seed=100).items()}
lm = model.language_model
L = cfg["text_config"]["num_hidden_layers"]
lm._trace = torch.zeros(1024 + 3 * (7 * L + 3), device="cuda", dtype=torch.int64)
lm._trace_cta = int(os.environ.get("TRACE_CTA", 0))
toks = model.generate(inp["input_ids"], inp["pixel_values"], inp["attention_mask"], 8, use_cuda_graph=False)
torch.cuda.synchronize()
t = lm._trace.cpu().numpy().astype("float64")
d = (t[1:7 * L + 2] - t[:7 * L + 1]) / 1e3
names = ["norm1", "qkv", "attn", "o", "norm2", "gate-up", "down"]
import numpy as np
per = d[:7 * L].reshape(L, 7)
print("phase durations (us), mean over layers 1..L-1, layer 0 separately")
for i, n in enumerate(names):
    print(f"  {n:8s} mean {per[1:, i].mean():7.2f}  min {per[1:, i].min():7.2f}  max {per[1:, i].max():7.2f}   layer0 {per[0, i]:7.2f}")
print(f"  final norm {d[7 * L]:7.2f}")
print(f"  layer total mean {per[1:].sum(1).mean():7.2f} us; all layers {per.sum():8.1f} us; step to final-norm barrier {(t[7 * L + 1] - t[0]) / 1e3:8.1f} us")

c = t[1024:].reshape(-1, 3)[1:7 * L + 2]   # rows: barrier index 1..; cols: work_done, local_done, grid_done
prev_done = np.concatenate([[np.nan], c[:-1, 2]])
work = (c[:, 0] - prev_done) / 1.9e3; local = (c[:, 1] - c[:, 0]) / 1.9e3; grid = (c[:, 2] - c[:, 1]) / 1.9e3
w = work[:7 * L].reshape(L, 7); lo = local[:7 * L].reshape(L, 7); g = grid[:7 * L].reshape(L, 7)
print("SM-clock detail for CTA", lm._trace_cta, "(us @1.9GHz): thread0 work | wait for CTA workers | grid barrier wait")
for i, n in enumerate(names):
    print(f"  {n:8s} work {np.nanmean(w[1:, i]):6.2f}  cta-sync {lo[1:, i].mean():6.2f}  grid {g[1:, i].mean():6.2f}")
