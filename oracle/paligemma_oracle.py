"""CPU fp32 ORACLE for the PaliGemma inference hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (CPU, fp32) restatement of the algorithm of prtk1729/Paligemma-MultiModal-System for the path
PaliGemmaForConditionalGeneration.forward + KVCache + the inference.py sampling loop.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` legs may import this module; the product package never does.

Parity pin: the reference has NO tests / golden vectors for this path, so this restatement is pinned
against outputs of the UNMODIFIED reference itself, executed in the build container by tests/golden/make_golden.py
(fixtures committed under tests/golden/, checked by tests/test_oracle_golden.py on every run).

It is written functionally over a state dict that uses the reference's parameter names.  Each function cites the
reference lines it follows.  Deliberate, semantics-preserving differences from the reference's execution:
  * the vision tower is not re-run on decode steps (reference re-runs it and discards the result,
    modeling_paligemma.py:281-282; pass `rerun_vision=True` to time the literal behaviour);
  * `last_only=True` computes logits for the last position only (the loop uses nothing else, inference.py:59);
  * decode works for batch > 1 with per-row position ids (the reference's (1,B) position_ids, modeling_paligemma.py:
    189-191, is only valid for B = 1; for B = 1 both agree exactly).
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------------------
# SigLIP vision tower (modeling_siglip.py)
# ----------------------------------------------------------------------------------------------------------------------
def siglip_embeddings(sd, vc, pixel_values, prefix="vision_tower.model.embeddings."):
    """modeling_siglip.py:280-299: valid conv P x P stride P -> flatten(2) -> transpose -> + positional embeddings."""
    P = vc["patch_size"]
    x = F.conv2d(pixel_values, sd[prefix + "patch_embedding.weight"], sd[prefix + "patch_embedding.bias"], stride=P)
    x = x.flatten(2).transpose(1, 2)
    return x + sd[prefix + "positional_embeddings.weight"][None, : x.shape[1]]


def siglip_attention(sd, vc, x, prefix):
    """modeling_siglip.py:65-157: biased q/k/v projections, 1/sqrt(dh) scale AFTER QK^T, fp32 softmax, out_proj."""
    B, N, D = x.shape
    H = vc["num_attention_heads"]
    dh = D // H
    q = F.linear(x, sd[prefix + "query_proj.weight"], sd[prefix + "query_proj.bias"]).view(B, N, H, dh).transpose(1, 2)
    k = F.linear(x, sd[prefix + "key_proj.weight"], sd[prefix + "key_proj.bias"]).view(B, N, H, dh).transpose(1, 2)
    v = F.linear(x, sd[prefix + "value_proj.weight"], sd[prefix + "value_proj.bias"]).view(B, N, H, dh).transpose(1, 2)
    w = torch.matmul(q, k.transpose(2, 3)) * (1.0 / dh ** 0.5)
    w = torch.softmax(w, dim=-1, dtype=torch.float32)
    o = torch.matmul(w, v).transpose(1, 2).reshape(B, N, D)
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def siglip_mlp(sd, x, prefix):
    """modeling_siglip.py:181-186: fc2(gelu_tanh(fc1(x)))."""
    h = F.gelu(F.linear(x, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"]), approximate="tanh")
    return F.linear(h, sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"])


def siglip_forward(sd, vc, pixel_values, prefix="vision_tower.model."):
    """modeling_siglip.py:206-221,234-239,312-320: pre-LN residual blocks, then post_layernorm."""
    eps = vc.get("layer_norm_eps", 1e-6)
    D = vc["hidden_size"]
    x = siglip_embeddings(sd, vc, pixel_values, prefix + "embeddings.")
    for i in range(vc["num_hidden_layers"]):
        lp = f"{prefix}encoder.layers.{i}."
        h = F.layer_norm(x, (D,), sd[lp + "layer_norm1.weight"], sd[lp + "layer_norm1.bias"], eps)
        x = x + siglip_attention(sd, vc, h, lp + "self_attn.")
        h = F.layer_norm(x, (D,), sd[lp + "layer_norm2.weight"], sd[lp + "layer_norm2.bias"], eps)
        x = x + siglip_mlp(sd, h, lp + "mlp.")
    return F.layer_norm(x, (D,), sd[prefix + "post_layernorm.weight"], sd[prefix + "post_layernorm.bias"], eps)


# ----------------------------------------------------------------------------------------------------------------------
# merge + positions (modeling_paligemma.py)
# ----------------------------------------------------------------------------------------------------------------------
def merge_embeddings(config, input_ids, text_embeds, image_feats):
    """modeling_paligemma.py:93-128: where(text) -> masked_scatter(image, feats * projection_dim^-0.5) -> where(pad, 0)."""
    pad = config.get("pad_token_id")
    pad = -1 if pad is None else pad
    img_tok = config["image_token_index"]
    D = text_embeds.shape[-1]
    is_pad = (input_ids == pad)[..., None].expand(-1, -1, D)
    is_img = (input_ids == img_tok)[..., None].expand(-1, -1, D)
    is_txt = ((input_ids != img_tok) & (input_ids != pad))[..., None].expand(-1, -1, D)
    out = torch.where(is_txt, text_embeds, torch.zeros_like(text_embeds))
    out = out.masked_scatter(is_img, image_feats * (config["projection_dim"] ** -0.5))
    return torch.where(is_pad, torch.zeros_like(out), out)


def position_ids(attention_mask, decode):
    """modeling_paligemma.py:187-195: prefill -> cumsum(mask) with 1 at masked slots; decode -> last cumsum, per row."""
    cs = attention_mask.cumsum(-1)
    if decode:
        return cs[:, -1:].long()
    return cs.masked_fill(attention_mask == 0, 1).long()


# ----------------------------------------------------------------------------------------------------------------------
# Gemma decoder (modeling_gemma.py)
# ----------------------------------------------------------------------------------------------------------------------
def rms_norm(x, w, eps=1e-6):
    """modeling_gemma.py:172-181 (eps is always the 1e-6 default: the layers never forward config.rms_norm_eps)."""
    x = x.float()
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * (1.0 + w.float())


def rope_cos_sin(pos, dh, theta=10000.0):
    """modeling_gemma.py:112,116-135: inv_freq = theta^(-2i/dh); angles fp32; emb = cat(freqs, freqs)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))
    freqs = pos[:, :, None].float() * inv_freq[None, None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def apply_rope(x, cos, sin):
    """modeling_gemma.py:138-151: x*cos + rotate_half(x)*sin, rotate_half = cat(-x2, x1); x is [B,H,S,dh]."""
    half = x.shape[-1] // 2
    rot = torch.cat((-x[..., half:], x[..., :half]), dim=-1)
    return x * cos[:, None] + rot * sin[:, None]


def gemma_layer(sd, tc, x, pos, kv, layer, prefix):
    """modeling_gemma.py:385-418 (DecoderLayer) with GemmaAttention :264-358 and GemmaMLP :210-218.
    `kv` is a list of [k, v] per layer (the KVCache of :8-64: append on first use, cat along dim -2 afterwards)."""
    B, S, D = x.shape
    Hq, Hkv, dh = tc["num_attention_heads"], tc["num_key_value_heads"], tc.get("head_dim", 256)
    h = rms_norm(x, sd[prefix + "input_layernorm.weight"])
    q = F.linear(h, sd[prefix + "self_attn.q_proj.weight"]).view(B, S, Hq, dh).transpose(1, 2)
    k = F.linear(h, sd[prefix + "self_attn.k_proj.weight"]).view(B, S, Hkv, dh).transpose(1, 2)
    v = F.linear(h, sd[prefix + "self_attn.v_proj.weight"]).view(B, S, Hkv, dh).transpose(1, 2)
    cos, sin = rope_cos_sin(pos, dh, tc.get("rope_theta", 10000.0))
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    if kv is not None:
        if layer >= len(kv):
            kv.append([k, v])
        else:
            kv[layer] = [torch.cat([kv[layer][0], k], dim=-2), torch.cat([kv[layer][1], v], dim=-2)]
        k, v = kv[layer]
    g = Hq // Hkv
    kk = k[:, :, None].expand(B, Hkv, g, k.shape[-2], dh).reshape(B, Hq, k.shape[-2], dh)
    vv = v[:, :, None].expand(B, Hkv, g, v.shape[-2], dh).reshape(B, Hq, v.shape[-2], dh)
    # the additive mask is all zeros in both phases (modeling_paligemma.py:154-169): full attention, padding unmasked
    w = torch.matmul(q, kk.transpose(-2, -1)) / math.sqrt(dh)
    w = torch.softmax(w, dim=-1, dtype=torch.float32)
    o = torch.matmul(w, vv).transpose(1, 2).reshape(B, S, Hq * dh)
    x = x + F.linear(o, sd[prefix + "self_attn.o_proj.weight"])
    h = rms_norm(x, sd[prefix + "post_attention_layernorm.weight"])
    m = F.gelu(F.linear(h, sd[prefix + "mlp.gate_proj.weight"]), approximate="tanh") * F.linear(h, sd[prefix + "mlp.up_proj.weight"])
    return x + F.linear(m, sd[prefix + "mlp.down_proj.weight"])


def gemma_forward(sd, tc, embeds, pos, kv, last_only=False, prefix="language_model."):
    """modeling_gemma.py:501-533: embeds * sqrt(D) -> layers -> final norm -> biased lm_head -> fp32 logits."""
    D = tc["hidden_size"]
    x = embeds * torch.tensor(D ** 0.5, dtype=embeds.dtype)
    for i in range(tc["num_hidden_layers"]):
        x = gemma_layer(sd, tc, x, pos, kv, i, f"{prefix}model.layers.{i}.")
    x = rms_norm(x, sd[prefix + "model.norm.weight"])
    if last_only:
        x = x[:, -1:, :]
    return F.linear(x, sd[prefix + "lm_head.weight"], sd[prefix + "lm_head.bias"]).float()


# ----------------------------------------------------------------------------------------------------------------------
# top level (modeling_paligemma.py:257-307)
# ----------------------------------------------------------------------------------------------------------------------
def image_features(sd, config, pixel_values):
    """vision tower + bias-free projector (modeling_paligemma.py:60-65,281-282)."""
    feats = siglip_forward(sd, config["vision_config"], pixel_values)
    return F.linear(feats, sd["multi_modal_projector.linear.weight"])


@torch.no_grad()
def forward(sd, config, input_ids, pixel_values, attention_mask, kv, last_only=False, image_feats=None, rerun_vision=False):
    """Returns fp32 logits [B, S or 1, V]; `kv` (list) is grown in place like the reference KVCache."""
    tc = config["text_config"]
    decode = kv is not None and len(kv) > 0
    if image_feats is None or rerun_vision:
        image_feats = image_features(sd, config, pixel_values)
    text = sd["language_model.model.embed_tokens.weight"][input_ids]
    if decode:
        assert input_ids.shape[1] == 1  # modeling_paligemma.py:161
        is_img = input_ids == config["image_token_index"]
        if bool(is_img.any()):  # masked_scatter with one True slot per row takes that image's first feature row
            feats = torch.stack([image_feats[b, :1] if is_img[b, 0] else torch.zeros_like(image_feats[b, :1])
                                 for b in range(input_ids.shape[0])])
            x = merge_embeddings(config, input_ids, text, feats)
        else:
            x = merge_embeddings(config, input_ids, text, image_feats[:, :0])
    else:
        x = merge_embeddings(config, input_ids, text, image_feats)
    pos = position_ids(attention_mask, decode)
    return gemma_forward(sd, tc, x, pos, kv, last_only=last_only)


def top_p_filter(probs, p):
    """inference.py:90-102: sort desc, drop where (cumsum - prob) > p, renormalise.  Returns (sorted probs, indices)."""
    srt, idx = torch.sort(probs, dim=-1, descending=True)
    cs = torch.cumsum(srt, dim=-1)
    srt = srt.masked_fill(cs - srt > p, 0.0)
    return srt / srt.sum(dim=-1, keepdim=True), idx


def sample_top_p(probs, p, generator=None):
    """inference.py:90-106."""
    srt, idx = top_p_filter(probs, p)
    return torch.gather(idx, -1, torch.multinomial(srt, 1, generator=generator))


@torch.no_grad()
def generate(sd, config, input_ids, pixel_values, attention_mask, max_new_tokens, do_sample=False, temperature=0.8,
             top_p=0.9, eos_token_id=None, generator=None, rerun_vision=False, return_logits=False, forced_tokens=None):
    """The loop of inference.py:45-79 for a batch (rows never stop early unless every row hit EOS; B = 1 matches the
    reference exactly, including appending EOS before the break).  `forced_tokens` [B, T] teacher-forces the inputs."""
    kv = []
    feats = image_features(sd, config, pixel_values)
    ids, mask = input_ids, attention_mask
    tokens, logits_log = [], []
    for step in range(max_new_tokens):
        logits = forward(sd, config, ids, pixel_values, mask, kv, last_only=True, image_feats=feats, rerun_vision=rerun_vision)[:, -1]
        if return_logits:
            logits_log.append(logits)
        if do_sample:
            nxt = sample_top_p(torch.softmax(logits / temperature, dim=-1), top_p, generator)
        else:
            nxt = torch.argmax(logits, dim=-1, keepdim=True)
        tokens.append(nxt)
        if eos_token_id is not None and bool((nxt == eos_token_id).all()):
            break
        ids = nxt if forced_tokens is None else forced_tokens[:, step: step + 1]
        mask = torch.cat([mask, torch.ones((mask.shape[0], 1), dtype=mask.dtype)], dim=-1)
    out = torch.cat(tokens, dim=-1)
    return (out, torch.stack(logits_log, 1)) if return_logits else out
