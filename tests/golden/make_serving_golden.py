"""Generates tests/golden/serving_reference.npz by executing the UNMODIFIED reference (/root/reference) on CPU fp32.

Run in the build container only:  python tests/golden/make_serving_golden.py
The reference serves one request at a time (inference.py:69), so every request of the ragged stream
(random_init.make_requests) is run through its loop on its own, unpadded: these are the answers a continuous batcher must
reproduce row by row (tests/test_serving_gpu.py) and the oracle must reproduce on CPU (tests/test_oracle_golden.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import build_reference, run_reference_loop  # noqa: E402  (also puts /root/reference on sys.path)

from paligemma_multimodal_system_b200.random_init import TINY_CONFIG, make_requests, make_state_dict  # noqa: E402

N_REQ, STEPS = 10, 12


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    out = {}
    reqs = make_requests(TINY_CONFIG, N_REQ, 2, 8, seed=21)
    out["prompt_lens"] = np.array([int(ids.numel()) for ids, _ in reqs])
    for regime in ("R0", "R1", "R2"):
        model = build_reference(TINY_CONFIG, make_state_dict(TINY_CONFIG, regime, seed=11))
        toks, first_logits = [], []
        for ids, px in reqs:
            inputs = dict(input_ids=ids[None], attention_mask=torch.ones(1, ids.numel(), dtype=torch.int64), pixel_values=px[None])
            t, l = run_reference_loop(model, inputs, STEPS)
            toks.append(t)
            first_logits.append(l[0])
        out[f"{regime}_tokens"] = np.stack(toks)
        out[f"{regime}_prefill_logits"] = np.stack(first_logits)
    path = os.path.join(HERE, "serving_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k, v in out.items():
        print(k, v.shape, v.dtype, v[:3].tolist() if "tokens" in k or "lens" in k else "")


if __name__ == "__main__":
    main()
