"""Continuous batching vs static batching on a ragged request stream (3B-224, one B200): bench.measure_serving with knobs.

    python scratch/serve_bench.py [--requests 256] [--slots 64] [--stage 32] [--min-admit 16] [--greedy] [--no-static]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_gpu_model, measure_serving  # noqa: E402
from paligemma_multimodal_system_b200.random_init import paligemma_3b_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--requests", type=int, default=256)
    ap.add_argument("--slots", type=int, default=64)
    ap.add_argument("--min-admit", type=int, default=16)
    ap.add_argument("--steps-per-replay", type=int, default=8)
    ap.add_argument("--stage", type=int, default=32)
    ap.add_argument("--greedy", action="store_true")
    ap.add_argument("--no-static", action="store_true")
    args = ap.parse_args()
    cfg = paligemma_3b_config(224)
    model, _ = build_gpu_model(cfg, seed=0)
    print(json.dumps(measure_serving(model, cfg, args.requests, args.slots, args.stage, args.min_admit, args.steps_per_replay,
                                     args.greedy, not args.no_static)))


if __name__ == "__main__":
    main()
