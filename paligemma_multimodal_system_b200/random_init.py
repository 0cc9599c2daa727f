"""Seeded random-init weights and synthetic inputs for PaliGemma (there is no checkpoint or dataset offline).

`make_state_dict` returns tensors under the reference's state-dict key names (the module tree of
modeling_paligemma.py:71-90 / modeling_siglip.py / modeling_gemma.py), every value rounded to a bf16-representable
fp32 number so that the fp32 CPU reference, the oracle and the bf16 GPU model hold IDENTICAL weights.

Init regimes:
  R0  "literal"  : PyTorch default init distributions (Linear/Conv2d U(+-1/sqrt(fan_in)), Embedding N(0,1),
                   LayerNorm 1/0, GemmaRMSNorm 0 as in modeling_gemma.py:163)
  R1  "wellcond" : tied embedding N(0, 0.02), LM matrices N(0, 1/sqrt(fan_in)), vision tower as R0
  R2  "diffuse"  : as R1 with LM matrices N(0, g/sqrt(fan_in)), g = 1.5 (large top-p kept set; throughput runs)
"""
import math

import torch

TINY_CONFIG = dict(
    vision_config=dict(hidden_size=256, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                       patch_size=14, image_size=224, num_channels=3, layer_norm_eps=1e-6),
    text_config=dict(hidden_size=256, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                     num_key_value_heads=1, head_dim=64, vocab_size=1281, rope_theta=10000.0, rms_norm_eps=1e-6),
    projection_dim=256, image_token_index=1024, pad_token_id=0, vocab_size=1281, hidden_size=256,
)


def paligemma_3b_config(image_size: int = 224) -> dict:
    """PaliGemma-3B-pt-{224,448,896}: SigLIP-So400m/14 + Gemma-2B (values of the HF config.json the reference cites at
    modeling_paligemma.py:8-9)."""
    return dict(
        vision_config=dict(hidden_size=1152, intermediate_size=4304, num_hidden_layers=27, num_attention_heads=16,
                           patch_size=14, image_size=image_size, num_channels=3, layer_norm_eps=1e-6),
        text_config=dict(hidden_size=2048, intermediate_size=16384, num_hidden_layers=18, num_attention_heads=8,
                         num_key_value_heads=1, head_dim=256, vocab_size=257216, rope_theta=10000.0, rms_norm_eps=1e-6),
        projection_dim=2048, image_token_index=257152, pad_token_id=0, vocab_size=257216, hidden_size=2048,
    )


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def make_state_dict(config: dict, regime: str = "R1", seed: int = 0, device="cpu", gain: float = 1.5,
                    dtype=torch.float32):
    """Reference-named state dict; values are bf16-representable; `dtype` is the storage dtype."""
    vc, tc = config["vision_config"], config["text_config"]
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}

    def uniform(shape, bound):
        return _bf16_round((torch.rand(shape, generator=g, device=device) * 2 - 1) * bound).to(dtype)

    def normal(shape, std):
        return _bf16_round(torch.randn(shape, generator=g, device=device) * std).to(dtype)

    def linear(prefix, out_f, in_f, bias, lm=False):
        if lm and regime in ("R1", "R2"):
            std = (gain if regime == "R2" else 1.0) / math.sqrt(in_f)
            sd[prefix + ".weight"] = normal((out_f, in_f), std)
        else:
            sd[prefix + ".weight"] = uniform((out_f, in_f), 1.0 / math.sqrt(in_f))
        if bias:
            sd[prefix + ".bias"] = uniform((out_f,), 1.0 / math.sqrt(in_f))

    Dv, Fv, P, C = vc["hidden_size"], vc["intermediate_size"], vc["patch_size"], vc.get("num_channels", 3)
    N = (vc["image_size"] // P) ** 2
    vp = "vision_tower.model."
    fan = C * P * P
    sd[vp + "embeddings.patch_embedding.weight"] = uniform((Dv, C, P, P), 1.0 / math.sqrt(fan))
    sd[vp + "embeddings.patch_embedding.bias"] = uniform((Dv,), 1.0 / math.sqrt(fan))
    sd[vp + "embeddings.positional_embeddings.weight"] = normal((N, Dv), 1.0)
    for i in range(vc["num_hidden_layers"]):
        lp = f"{vp}encoder.layers.{i}."
        for ln in ("layer_norm1", "layer_norm2"):
            sd[lp + ln + ".weight"] = torch.ones(Dv, device=device, dtype=dtype)
            sd[lp + ln + ".bias"] = torch.zeros(Dv, device=device, dtype=dtype)
        for proj in ("key_proj", "value_proj", "query_proj", "out_proj"):
            linear(lp + "self_attn." + proj, Dv, Dv, True)
        linear(lp + "mlp.fc1", Fv, Dv, True)
        linear(lp + "mlp.fc2", Dv, Fv, True)
    sd[vp + "post_layernorm.weight"] = torch.ones(Dv, device=device, dtype=dtype)
    sd[vp + "post_layernorm.bias"] = torch.zeros(Dv, device=device, dtype=dtype)

    D, F, V = tc["hidden_size"], tc["intermediate_size"], tc["vocab_size"]
    Hq, Hkv, dh = tc["num_attention_heads"], tc["num_key_value_heads"], tc.get("head_dim", 256)
    linear("multi_modal_projector.linear", config.get("projection_dim", 2048), Dv, False)
    lm = "language_model."
    emb = normal((V, D), 1.0 if regime == "R0" else 0.02)
    pad = config.get("pad_token_id")
    if pad is not None and 0 <= pad < V:
        emb[pad].zero_()  # nn.Embedding(padding_idx=...) zero-initialises that row (modeling_gemma.py:437-439)
    sd[lm + "model.embed_tokens.weight"] = emb
    sd[lm + "lm_head.weight"] = emb  # tied (modeling_gemma.py:492-499)
    sd[lm + "lm_head.bias"] = uniform((V,), 1.0 / math.sqrt(D))
    for i in range(tc["num_hidden_layers"]):
        lp = f"{lm}model.layers.{i}."
        sd[lp + "input_layernorm.weight"] = torch.zeros(D, device=device, dtype=dtype)
        sd[lp + "post_attention_layernorm.weight"] = torch.zeros(D, device=device, dtype=dtype)
        linear(lp + "self_attn.q_proj", Hq * dh, D, False, lm=True)
        linear(lp + "self_attn.k_proj", Hkv * dh, D, False, lm=True)
        linear(lp + "self_attn.v_proj", Hkv * dh, D, False, lm=True)
        linear(lp + "self_attn.o_proj", D, D, False, lm=True)
        linear(lp + "mlp.gate_proj", F, D, False, lm=True)
        linear(lp + "mlp.up_proj", F, D, False, lm=True)
        linear(lp + "mlp.down_proj", D, F, False, lm=True)
    sd[lm + "model.norm.weight"] = torch.zeros(D, device=device, dtype=dtype)
    return sd


def make_inputs(config: dict, batch: int, prompt_len: int = 4, seed: int = 0, device="cpu"):
    """Synthetic request batch: `<image>`*N + bos(2) + random prompt tokens + "\\n"(108); all-ones
    attention mask; pixel values uniform in [-1, 1], bf16-representable.  Row r is seeded with seed + r."""
    vc, tc = config["vision_config"], config["text_config"]
    N = (vc["image_size"] // vc["patch_size"]) ** 2
    V_text = min(config["image_token_index"], tc["vocab_size"])
    ids, px = [], []
    for r in range(batch):
        g = torch.Generator(device="cpu").manual_seed(seed * 1000003 + r)
        mid = torch.randint(3, V_text, (max(prompt_len - 2, 0),), generator=g)
        row = torch.cat([torch.full((N,), config["image_token_index"]), torch.tensor([2]), mid,
                         torch.tensor([108 if V_text > 108 else 3])]).long()
        ids.append(row)
        px.append(_bf16_round(torch.rand(vc.get("num_channels", 3), vc["image_size"], vc["image_size"], generator=g) * 2 - 1))
    input_ids = torch.stack(ids).to(device)
    return dict(input_ids=input_ids, attention_mask=torch.ones_like(input_ids), pixel_values=torch.stack(px).to(device))


def make_requests(config: dict, n: int, min_prompt_len: int = 2, max_prompt_len: int = 8, seed: int = 0):
    """Synthetic request stream with RAGGED prompts (serving.py): request r = `<image>`*N + bos(2) + random tokens, text
    length cycling through [min_prompt_len, max_prompt_len], the last token random per request (so that even the
    degenerate random-init regimes, which echo the last prompt token, give every request its own continuation); pixel
    values as in make_inputs.  Returns a list of (input_ids int64 [S_r], pixel_values [3, H, W])."""
    vc, tc = config["vision_config"], config["text_config"]
    N = (vc["image_size"] // vc["patch_size"]) ** 2
    V_text = min(config["image_token_index"], tc["vocab_size"])
    out = []
    span = max_prompt_len - min_prompt_len + 1
    for r in range(n):
        g = torch.Generator(device="cpu").manual_seed(seed * 1000003 + 7919 * r + 1)
        t = min_prompt_len + (r * 5) % span
        text = torch.cat([torch.tensor([2]), torch.randint(3, V_text, (t - 1,), generator=g)]).long()
        ids = torch.cat([torch.full((N,), config["image_token_index"]).long(), text])
        px = _bf16_round(torch.rand(vc.get("num_channels", 3), vc["image_size"], vc["image_size"], generator=g) * 2 - 1)
        out.append((ids, px))
    return out
