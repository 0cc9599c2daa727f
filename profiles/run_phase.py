"""Profiling driver: builds the 3B-224 model, warms up, then brackets ONE phase with cudaProfilerStart/Stop so that
`ncu --profile-from-start off` sees only that phase.

    python profiles/run_phase.py --phase decode --batch 64 --steps 2     # eager decode steps (kernel by kernel)
    python profiles/run_phase.py --phase prefill --batch 8
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import PROMPT_LEN, build_gpu_model  # noqa: E402
from paligemma_multimodal_system_b200.random_init import make_inputs, paligemma_3b_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--phase", default="decode", choices=["decode", "prefill"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--image-size", type=int, default=224)
    ap.add_argument("--greedy", action="store_true")
    a = ap.parse_args()
    cfg = paligemma_3b_config(a.image_size)
    model, _ = build_gpu_model(cfg)
    inp = {k: v.cuda() for k, v in make_inputs(cfg, batch=a.batch, prompt_len=PROMPT_LEN, seed=100).items()}
    gen = dict(do_sample=not a.greedy, temperature=0.8, top_p=0.9, seed=1234, use_cuda_graph=False)
    model.generate(inp["input_ids"], inp["pixel_values"], inp["attention_mask"], 3, **gen)  # warm-up (lazy init, attributes)
    torch.cuda.synchronize()
    if a.phase == "prefill":
        torch.cuda.profiler.start()
        model.forward(inp["input_ids"], inp["pixel_values"], inp["attention_mask"], None, last_only=True)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    else:
        # generate(): token 0 from prefill, then `steps` eager decode steps; start the profiler after the prefill
        orig = model.language_model.prefill

        def prefill_then_start(*args, **kw):
            out = orig(*args, **kw)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            return out

        model.language_model.prefill = prefill_then_start
        model.generate(inp["input_ids"], inp["pixel_values"], inp["attention_mask"], a.steps + 1, **gen)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("done")


if __name__ == "__main__":
    main()
