"""Where does a real decode step go?  Times CUDA-graph replays of sub-sequences of the real 3B model's step."""
import os, sys, torch
sys.path.insert(0, '.')
from bench import PROMPT_LEN, build_gpu_model
from paligemma_multimodal_system_b200 import _lib
from paligemma_multimodal_system_b200.random_init import make_inputs, paligemma_3b_config
cfg = paligemma_3b_config(224)
model, _ = build_gpu_model(cfg)
B = 64
inp = {k: v.cuda() for k, v in make_inputs(cfg, batch=B, prompt_len=PROMPT_LEN, seed=100).items()}
toks = model.generate(inp["input_ids"], inp["pixel_values"], inp["attention_mask"], 66, do_sample=False)
torch.cuda.synchronize()
lm = model.language_model
key = next(iter(model._graphs))
stt = model._graphs[key]
kv = stt["kv"]; cur = stt["cur"]; nxt = stt["nxt"]; hist = stt["hist"]; step = stt["step"]
bufs = lm.decode_buffers(B)
pk = lm._packed
L = _lib.lib()
V = cfg["text_config"]["vocab_size"]; D = cfg["text_config"]["hidden_size"]
def graph_time(name, fn, reps=20):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{name:50s} {us:9.2f} us", flush=True)
    return us
def full():
    model._decode_step(cur, kv, bufs, B)
    _lib.check(L.pg_argmax(bufs["logits"].data_ptr(), V, nxt.data_ptr(), B, V, _lib.stream()), "argmax")
def layers_only(n):
    def f():
        import math
        c = lm.text_config
        saved = pk["layers"]
        pk["layers"] = saved[:n]
        try:
            lm.decode_layers(bufs, kv, B)
        finally:
            pk["layers"] = saved
    return f
t_full = graph_time("full step (embed+18 layers+norm+head+argmax)", full)
graph_time("full step, 1500 replays (2 s sustained)", full, reps=1500)
graph_time("full step again, 20 replays", full)
t18 = graph_time("18 layers + norm + head", layers_only(18))
t9 = graph_time(" 9 layers + norm + head", layers_only(9))
t1 = graph_time(" 1 layer  + norm + head", layers_only(1))
print(f"per layer (18 vs 9): {(t18 - t9) / 9:.2f} us; (9 vs 1): {(t9 - t1) / 8:.2f} us; norm+head+graph overhead ~ {t1 - (t18 - t9) / 9:.2f} us")
for nm in ("h", "hn", "qkv", "att", "mid"):
    print(nm, bufs[nm].shape, bufs[nm].data_ptr() % 1024)
print("kv pages", kv.k_pages.shape, "max_pages", kv.page_table.shape[1], "kv_len", kv.counters[2][:4].tolist())
