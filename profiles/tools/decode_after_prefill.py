"""How the decode step time evolves after the prefill burst inside generate(): T = 9, 17, 33, 65, 129, 257 new tokens, 3B-224,
64 requests; decode_ms differences give the mean step time of each segment of the job.  Also samples SM clock / power."""
import os, sys, threading, time, subprocess, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from paligemma_multimodal_system_b200.random_init import make_inputs, paligemma_3b_config
cfg = paligemma_3b_config(224)
model, _ = bench.build_gpu_model(cfg)
gen = dict(do_sample=True, temperature=0.8, top_p=0.9, seed=1)
inp = make_inputs(cfg, batch=64, prompt_len=bench.PROMPT_LEN, seed=100)
dev = {k: v.cuda() for k, v in inp.items()}
prev_T, prev_ms = 1, 0.0
for T in (9, 17, 33, 65, 129, 257):
    for _ in range(2):
        model.generate(dev["input_ids"], dev["pixel_values"], dev["attention_mask"], T, **gen)
    torch.cuda.synchronize()
    ms = pre = 0.0
    R = 4
    for _ in range(R):
        tm = {}
        model.generate(dev["input_ids"], dev["pixel_values"], dev["attention_mask"], T, timings=tm, **gen)
        ms += tm["decode_ms"] / R; pre += tm["prefill_ms"] / R
    print(f"T={T:4d}: prefill {pre:6.1f} ms, decode {ms:7.2f} ms = {ms / (T - 1):.4f} ms/step; steps {prev_T}..{T - 1}: {(ms - prev_ms) / (T - prev_T):.4f} ms/step", flush=True)
    prev_T, prev_ms = T, ms
# idle gap between prefill and decode: does a pause let the clocks come back?
