"""Continuous batching vs static batching on a ragged request stream (3B-224, one B200).

    python scratch/serve_bench.py [--requests 256] [--slots 64] [--min-admit 8] [--greedy]

Stream: `requests` synthetic requests, text prompts of 2..12 tokens, token budgets drawn uniformly from [16, 128] (the
stand-in for EOS-terminated answers of different lengths; random-init weights never favour EOS).  Static arm: the stream
in arrival order, 64 at a time through generate() (prompts padded per batch to... not possible in the reference API:
generate() takes a dense batch, so the static arm uses the longest budget of each batch and one prompt length -- it is the
upper bound on static batching with perfect length bucketing).  Useful tokens = sum of budgets in both arms."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_gpu_model  # noqa: E402
from paligemma_multimodal_system_b200.random_init import make_inputs, make_requests, paligemma_3b_config  # noqa: E402
from paligemma_multimodal_system_b200.serving import ContinuousBatcher  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--requests", type=int, default=256)
    ap.add_argument("--slots", type=int, default=64)
    ap.add_argument("--min-admit", type=int, default=8)
    ap.add_argument("--steps-per-replay", type=int, default=8)
    ap.add_argument("--stage", type=int, default=16)
    ap.add_argument("--greedy", action="store_true")
    args = ap.parse_args()
    cfg = paligemma_3b_config(224)
    model, _ = build_gpu_model(cfg, seed=0)
    n = args.requests
    reqs = make_requests(cfg, n, 2, 12, seed=3)
    g = torch.Generator().manual_seed(5)
    budgets = torch.randint(16, 129, (n,), generator=g).tolist()
    useful = sum(budgets)
    gen = dict(do_sample=not args.greedy, temperature=0.8, top_p=0.9, seed=1)

    def run_cb():
        cb = ContinuousBatcher(model, num_slots=args.slots, max_prompt_len=256 + 12, max_new_tokens=128, min_admit=args.min_admit,
                               steps_per_replay=args.steps_per_replay, stage=args.stage, **gen)
        for rep in range(2):  # first wave: warm-up + graph capture
            for (ids, px), m in zip(reqs, budgets):
                cb.submit(ids, px, m)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = cb.run()
            dt = time.perf_counter() - t0
            st = dict(cb.stats)
            cb.reset_stats()
        assert sum(len(v) for v in out.values()) == useful
        return dt, st

    def run_static():
        B = args.slots
        inp = make_inputs(cfg, batch=B, prompt_len=7, seed=9)
        dev = {k: v.cuda() for k, v in inp.items()}
        dt = 0.0
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for lo in range(0, n, B):
                T = max(budgets[lo:lo + B])
                model.generate(dev["input_ids"], dev["pixel_values"], dev["attention_mask"], T, **gen)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        return dt

    dt_s = run_static()
    dt_c, st = run_cb()
    print(f"requests {n}, slots {args.slots}, stage {args.stage}, min_admit {args.min_admit}, useful tokens {useful}, budgets 16..128, prompts 258..268 tokens")
    print(f"static batching   : {dt_s * 1e3:8.1f} ms  {useful / dt_s:9.0f} useful tok/s  ({n / dt_s:.1f} requests/s)")
    print(f"continuous batching: {dt_c * 1e3:8.1f} ms  {useful / dt_c:9.0f} useful tok/s  ({n / dt_c:.1f} requests/s)  "
          f"[{st['prefill_groups']} prefill groups, {st['decode_steps']} decode steps, slot occupancy "
          f"{useful / max(1, st['decode_steps'] * args.slots):.2f}]")


if __name__ == "__main__":
    main()
