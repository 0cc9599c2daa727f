"""Shared helpers for the GPU parity tests (oracle = CPU fp32 restatement pinned to the reference)."""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build_model(config, sd, dtype=torch.bfloat16):
    from paligemma_multimodal_system_b200.modeling_paligemma import PaliGemmaConfig, PaliGemmaForConditionalGeneration
    model = PaliGemmaForConditionalGeneration(PaliGemmaConfig(**copy.deepcopy(config)), device="cuda", dtype=dtype).eval()
    model.load_state_dict(sd, strict=True)
    model.tie_weights()
    return model


def stats(got, ref):
    got, ref = got.detach().float().cpu().flatten(), ref.detach().float().cpu().flatten()
    err = (got - ref).abs().max().item()
    absmax = ref.abs().max().item()
    cos = torch.nn.functional.cosine_similarity(got.double(), ref.double(), dim=0).item()
    return dict(max_abs=err, rel=err / max(absmax, 1e-12), cos=cos, absmax=absmax)


def top2_margin(logits):
    v = torch.topk(logits.float(), 2, dim=-1).values
    return (v[..., 0] - v[..., 1])
