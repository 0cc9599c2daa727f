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
# replay the captured graphs directly, timing each replay
evs = [torch.cuda.Event(enable_timing=True) for _ in range(40)]
tm = {}
model.generate(inp["input_ids"], inp["pixel_values"], inp["attention_mask"], T, do_sample=False, timings=tm)
print("generate timings", tm, "per step", tm["decode_ms"] / 127 * 1e3)
# now: state after a generate (kv_len ~ 388).  Reset counters to the start and replay the K-graph 15 times with events
kv = stt["kv"]
kv.counters[0].fill_(261); kv.counters[1].fill_(260); kv.counters[2].fill_(261); stt["step"].zero_()
torch.cuda.synchronize()
evs[0].record()
for i in range(15):
    stt["graph_k"].replay()
    evs[i + 1].record()
torch.cuda.synchronize()
print("per 8-step replay (us/step):", [round(evs[i].elapsed_time(evs[i + 1]) * 1e3 / 8, 1) for i in range(15)])
print("kv_len now", kv.counters[2][:2].tolist())
# does the prefill burst (1 kW tensor-core phase) slow the decode replays that follow it?
for trial in range(2):
    kv.counters[0].fill_(261); kv.counters[1].fill_(260); kv.counters[2].fill_(261); stt["step"].zero_()
    h, pos = model._merge(inp["input_ids"], inp["attention_mask"], stt["img"])
    model.language_model.prefill(h, pos, B, 260, kv, last_only=True)
    kv.counters[0].fill_(261); kv.counters[1].fill_(260); kv.counters[2].fill_(261); stt["step"].zero_()
    evs[0].record()
    for i in range(15):
        stt["graph_k"].replay()
        evs[i + 1].record()
    torch.cuda.synchronize()
    print("after prefill, per 8-step replay (us/step):", [round(evs[i].elapsed_time(evs[i + 1]) * 1e3 / 8, 1) for i in range(15)])
