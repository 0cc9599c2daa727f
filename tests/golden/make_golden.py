"""Generates tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference) on CPU fp32.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Weights and inputs come from paligemma_multimodal_system_b200.random_init (seeded, bf16-representable), loaded into the
reference modules with load_state_dict(strict=True) + tie_weights(), so tests can rebuild the identical state dict.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("fire", types.SimpleNamespace(Fire=lambda f: None))  # inference.py imports fire (absent here)

from paligemma_multimodal_system_b200.random_init import TINY_CONFIG, make_inputs, make_state_dict  # noqa: E402

import inference as ref_inference  # noqa: E402
from modeling_gemma import KVCache  # noqa: E402
from modeling_paligemma import PaliGemmaConfig, PaliGemmaForConditionalGeneration  # noqa: E402


def build_reference(config, sd):
    import copy
    model = PaliGemmaForConditionalGeneration(PaliGemmaConfig(**copy.deepcopy(config))).eval()
    missing, unexpected = model.load_state_dict(sd, strict=True)
    model.tie_weights()
    return model


@torch.no_grad()
def run_reference_loop(model, inputs, steps, do_sample=False, temperature=0.8, top_p=0.9):
    """inference.py:45-79 driven at model level (no tokenizer / image file offline); B = 1."""
    ids, mask, px = inputs["input_ids"], inputs["attention_mask"], inputs["pixel_values"]
    kv = KVCache()
    toks, logs = [], []
    for _ in range(steps):
        with contextlib.redirect_stdout(io.StringIO()):
            out = model(input_ids=ids, pixel_values=px, attention_mask=mask, kv_cache=kv)
        kv = out["kv_cache"]
        logits = out["logits"][:, -1, :]
        logs.append(logits[0].clone())
        if do_sample:
            nxt = ref_inference._sample_top_p(torch.softmax(logits / temperature, dim=-1), top_p)
        else:
            nxt = torch.argmax(logits, dim=-1, keepdim=True)
        assert nxt.size() == (1, 1)
        toks.append(int(nxt))
        ids = nxt.squeeze(0).unsqueeze(-1)
        mask = torch.cat([mask, torch.ones((1, 1))], dim=-1)
    return np.array(toks, dtype=np.int64), torch.stack(logs).numpy()


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    out = {}
    for regime in ("R0", "R1", "R2"):
        sd = make_state_dict(TINY_CONFIG, regime, seed=11)
        model = build_reference(TINY_CONFIG, sd)
        inputs = make_inputs(TINY_CONFIG, batch=1, prompt_len=4, seed=5)
        toks, logs = run_reference_loop(model, inputs, 16)
        out[f"{regime}_greedy_tokens"] = toks
        out[f"{regime}_step_logits"] = logs
        torch.manual_seed(1234)
        stoks, _ = run_reference_loop(model, inputs, 16, do_sample=True)
        out[f"{regime}_sampled_tokens_seed1234"] = stoks
        # batched prefill (the reference supports B > 1 only here): full logits at 3 positions, vision/projector slices
        inputs2 = make_inputs(TINY_CONFIG, batch=2, prompt_len=6, seed=7)
        with contextlib.redirect_stdout(io.StringIO()):
            o2 = model(input_ids=inputs2["input_ids"], pixel_values=inputs2["pixel_values"],
                       attention_mask=inputs2["attention_mask"], kv_cache=KVCache())
            feats = model.vision_tower(inputs2["pixel_values"])
            proj = model.multi_modal_projector(feats)
        out[f"{regime}_b2_logits_pos"] = o2["logits"][:, [0, 255, -1], :].numpy()
        out[f"{regime}_b2_vision_slice"] = feats[:, :4, :32].numpy()
        out[f"{regime}_b2_vision_norm"] = feats.norm(dim=-1).numpy()
        out[f"{regime}_b2_proj_slice"] = proj[:, :4, :32].numpy()
        k0 = o2["kv_cache"].k_cache[1]
        out[f"{regime}_b2_kcache_l1_slice"] = k0[:, 0, -3:, :].numpy()
        out[f"{regime}_b2_num_items"] = np.array(o2["kv_cache"].num_items())
        # padded prompt (pad id 0, mask 0): the reference zeroes pad embeddings and gives them position 1
        ids3 = inputs2["input_ids"].clone()
        mask3 = inputs2["attention_mask"].clone()
        ids3[1, -1] = 0
        mask3[1, -1] = 0
        with contextlib.redirect_stdout(io.StringIO()):
            o3 = model(input_ids=ids3, pixel_values=inputs2["pixel_values"], attention_mask=mask3, kv_cache=KVCache())
        out[f"{regime}_b2_padded_logits_last"] = o3["logits"][:, -1, :].numpy()

    # _sample_top_p kept sets on synthetic probabilities (inference.py:90-106)
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(4, 1281, generator=g) * torch.tensor([0.5, 1.0, 2.0, 4.0])[:, None]
    probs = torch.softmax(logits / 0.8, dim=-1)
    srt, idx = torch.sort(probs, dim=-1, descending=True)
    cs = torch.cumsum(srt, dim=-1)
    kept = (~(cs - srt > 0.9)).sum(-1)
    out["topp_logits"] = logits.numpy()
    out["topp_kept_count"] = kept.numpy()
    torch.manual_seed(77)
    out["topp_samples_seed77"] = torch.cat([ref_inference._sample_top_p(probs, 0.9) for _ in range(8)], -1).numpy()

    # SigLIP smoke shape pinned by the reference's own __main__ (modeling_siglip.py:337-360)
    from modeling_siglip import SiglipVisionConfig, SiglipVisionModel
    torch.manual_seed(0)
    m = SiglipVisionModel(SiglipVisionConfig(num_channels=3, image_size=224, patch_size=16, hidden_size=768,
                                              intermediate_size=3072, num_hidden_layers=1, num_attention_heads=12))
    out["siglip_base_shape"] = np.array(m(torch.rand(2, 3, 224, 224)).shape)

    path = os.path.join(HERE, "tiny_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
