"""Does an idle gap after the 1 kW prefill burst let the SM clocks recover faster than decoding through the droop?"""
import os, sys, torch, time
sys.path.insert(0, '.')
from bench import PROMPT_LEN, build_gpu_model
from paligemma_multimodal_system_b200.random_init import make_inputs, paligemma_3b_config
cfg = paligemma_3b_config(224)
model, _ = build_gpu_model(cfg)
B, T = 64, 128
inp = {k: v.cuda() for k, v in make_inputs(cfg, batch=B, prompt_len=PROMPT_LEN, seed=100).items()}
for _ in range(2):
    model.generate(inp["input_ids"], inp["pixel_values"], inp["attention_mask"], T, do_sample=False)
torch.cuda.synchronize()
stt = model._graphs[next(iter(model._graphs))]
kv = stt["kv"]
def reset():
    kv.counters[0].fill_(261); kv.counters[1].fill_(260); kv.counters[2].fill_(261); stt["step"].zero_()
def run(gap_ms, mode):
    reset()
    img = model.image_features(inp["pixel_values"])
    h, pos = model._merge(inp["input_ids"], inp["attention_mask"], img)
    model.language_model.prefill(h, pos, B, 260, kv, last_only=True)
    reset()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    if gap_ms > 0:
        if mode == "sleep":
            torch.cuda._sleep(int(gap_ms * 1.9e6))
        else:  # host-side gap: the GPU is truly idle
            torch.cuda.synchronize(); time.sleep(gap_ms / 1e3)
    e1.record()
    for i in range(15): stt["graph_k"].replay()
    for i in range(7): stt["graph"].replay()
    e2.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), e1.elapsed_time(e2)
for mode in ("sleep", "host"):
    for gap in (0, 1, 3, 5, 10, 20, 40):
        res = [run(gap, mode) for _ in range(3)]
        g = sum(r[0] for r in res) / 3; d = sum(r[1] for r in res) / 3
        print(f"{mode:6s} gap {gap:3d} ms: gap {g:6.2f} ms + decode(127 steps) {d:7.2f} ms = {g + d:7.2f} ms   ({d / 127 * 1e3:.0f} us/step)", flush=True)
