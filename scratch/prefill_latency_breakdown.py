"""B = 1 prefill (cfg-2 latency path): is it bound by the host enqueue rate or by the kernels?"""
import sys, time, torch
sys.path.insert(0, '.')
from bench import PROMPT_LEN, build_gpu_model
from paligemma_multimodal_system_b200 import _lib
from paligemma_multimodal_system_b200.modeling_gemma import KVCache
from paligemma_multimodal_system_b200.random_init import make_inputs, paligemma_3b_config
cfg = paligemma_3b_config(224)
model, _ = build_gpu_model(cfg)
lm = model.language_model
for B in (1, 4):
    inp = {k: v.cuda() for k, v in make_inputs(cfg, batch=B, prompt_len=PROMPT_LEN, seed=100).items()}
    S = inp["input_ids"].shape[1]
    kv = KVCache(); kv.allocate(B, 18, 1, 256, S + 64)

    def vision():
        return model.image_features(inp["pixel_values"])

    def lang(img):
        h, pos = model._merge(inp["input_ids"], inp["attention_mask"], img)
        return lm.prefill(h, pos, B, S, kv, last_only=True)

    for _ in range(3):
        lang(vision())
    torch.cuda.synchronize()

    def timed(fn, n=5):
        out = []
        for _ in range(n):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0 = _lib.lib().pg_launch_count()
            t0 = time.perf_counter(); e0.record(); r = fn(); e1.record(); t1 = time.perf_counter()
            torch.cuda.synchronize()
            out.append((e0.elapsed_time(e1), (t1 - t0) * 1e3, _lib.lib().pg_launch_count() - c0))
        out.sort()
        return out[len(out) // 2], r

    (gv, hv, nv), img = timed(vision)
    (gl, hl, nl), _ = timed(lambda: lang(img))
    (ga, ha, na), _ = timed(lambda: lang(vision()))
    print(f"B={B}: vision gpu {gv:.2f} ms host-enqueue {hv:.2f} ms ({nv} launches) | merge+gemma gpu {gl:.2f} host {hl:.2f} ({nl}) | "
          f"both gpu {ga:.2f} host {ha:.2f}")
    # the same kernels replayed from a CUDA graph (no host in the loop): kernel-limited time
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    ids, mask = inp["input_ids"], inp["attention_mask"]
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            img2 = model.image_features(inp["pixel_values"])
            pk = lm._packed
            D = 2048
            h = torch.empty(B * S, D, device="cuda", dtype=torch.float32)
            pos = torch.empty(B * S, device="cuda", dtype=torch.int32)
            src = torch.empty(B * S, device="cuda", dtype=torch.int32)
            err = torch.zeros(1, device="cuda", dtype=torch.int32)
            _lib.check(_lib.lib().pg_merge_embeddings(ids.data_ptr(), mask.data_ptr(), pk["embed"].data_ptr(), img2.data_ptr(), h.data_ptr(),
                       pos.data_ptr(), src.data_ptr(), err.data_ptr(), B, S, D, 256, model.dummy_image_token_id, model.pad_token_id,
                       D ** 0.5, 1.0, _lib.stream()), "merge")
            lg = lm.prefill(h, pos, B, S, kv, last_only=True)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        g.replay()
    (gg, hg, _), _ = timed(lambda: g.replay())
    print(f"B={B}: CUDA-graph replay of the whole prefill: gpu {gg:.2f} ms (host {hg:.2f})")
