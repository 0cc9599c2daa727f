"""utils.load_hf_model (reference utils.py:9-37): a synthetic checkpoint under HUGGING FACE key names -> model.

There is no real checkpoint offline, so the test writes one: seeded random-init weights (random_init.make_state_dict),
renamed the way `PaliGemmaForConditionalGeneration.save_pretrained` names them (`vision_tower.vision_model...k_proj`,
`position_embedding`, no `lm_head.*` at all, a projector bias), sharded over two safetensors files, plus config.json."""
import copy
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from paligemma_multimodal_system_b200.random_init import TINY_CONFIG, make_inputs, make_state_dict  # noqa: E402

_REF_TO_HF = (
    ("vision_tower.model.", "vision_tower.vision_model."),
    (".embeddings.positional_embeddings.", ".embeddings.position_embedding."),
    (".self_attn.key_proj.", ".self_attn.k_proj."),
    (".self_attn.value_proj.", ".self_attn.v_proj."),
    (".self_attn.query_proj.", ".self_attn.q_proj."),
)


class _Tok:
    padding_side = "right"


def _write_checkpoint(path, sd, proj_bias=None, extra=None):
    from safetensors.torch import save_file
    hf = {}
    for k, v in sd.items():
        if k.startswith("language_model.lm_head."):
            continue  # tied weight, no bias: neither is stored in a PaliGemma checkpoint
        if k.startswith("vision_tower."):
            for a, b in _REF_TO_HF:
                k = k.replace(a, b)
        hf[k] = v.to(torch.bfloat16).contiguous()
    if proj_bias is not None:
        hf["multi_modal_projector.linear.bias"] = proj_bias.to(torch.bfloat16)
    hf["vision_tower.vision_model.embeddings.position_ids"] = torch.arange(256).view(1, -1)  # old-transformers buffer
    if extra:
        hf.update(extra)
    keys = sorted(hf)
    save_file({k: hf[k] for k in keys[::2]}, os.path.join(path, "model-00001-of-00002.safetensors"))
    save_file({k: hf[k] for k in keys[1::2]}, os.path.join(path, "model-00002-of-00002.safetensors"))
    cfg = copy.deepcopy(TINY_CONFIG)
    cfg.update(model_type="paligemma", architectures=["PaliGemmaForConditionalGeneration"])  # extra keys HF writes
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f)


def test_loader_accounts_for_every_key_cpu(tmp_path):
    """CPU: names are remapped, a tied / absent lm_head is fine, its bias is zeroed (the reference leaves it random), and a
    key that matches nothing raises instead of being dropped by strict=False."""
    from paligemma_multimodal_system_b200.utils import load_hf_model, read_hf_checkpoint
    sd = make_state_dict(TINY_CONFIG, "R1", seed=21)
    _write_checkpoint(str(tmp_path), sd)
    tensors, cfg = read_hf_checkpoint(str(tmp_path))
    assert "vision_tower.model.encoder.layers.1.self_attn.key_proj.weight" in tensors
    assert "vision_tower.model.embeddings.positional_embeddings.weight" in tensors
    model, tok = load_hf_model(str(tmp_path), device="cpu", tokenizer_loader=lambda p: _Tok())
    assert isinstance(tok, _Tok)
    got = model.state_dict()
    for k, v in sd.items():
        if k == "language_model.lm_head.bias":
            assert torch.count_nonzero(got[k]) == 0
        else:
            assert torch.equal(got[k].float(), v.float()), k
    assert model.language_model.lm_head.weight is model.language_model.model.embed_tokens.weight
    bad = tmp_path / "bad"
    bad.mkdir()
    _write_checkpoint(str(bad), sd, extra={"language_model.model.layers.0.self_attn.qkv_proj.weight": torch.zeros(4, 4)})
    with pytest.raises(KeyError, match="unexpected"):
        load_hf_model(str(bad), device="cpu", tokenizer_loader=lambda p: _Tok())
    short = tmp_path / "short"
    short.mkdir()
    _write_checkpoint(str(short), {k: v for k, v in sd.items() if "layers.1.mlp.down_proj" not in k})
    with pytest.raises(KeyError, match="missing"):
        load_hf_model(str(short), device="cpu", tokenizer_loader=lambda p: _Tok())


@pytest.mark.gpu
def test_load_hf_model_logits_match_direct_state_dict_and_oracle(tmp_path):
    """GPU: safetensors (HF names) -> load_hf_model -> packed device weights; prefill + 4 teacher-forced decode steps give the
    logits of the model built straight from the reference-named state dict, and of the CPU oracle."""
    from oracle import paligemma_oracle as O
    from paligemma_multimodal_system_b200.modeling_paligemma import PaliGemmaConfig, PaliGemmaForConditionalGeneration
    from paligemma_multimodal_system_b200.utils import load_hf_model
    sd = make_state_dict(TINY_CONFIG, "R1", seed=22)
    sd["language_model.lm_head.bias"].zero_()  # what the checkpoint means (no lm_head bias)
    _write_checkpoint(str(tmp_path), sd)
    loaded, _ = load_hf_model(str(tmp_path), device="cuda", tokenizer_loader=lambda p: _Tok())
    direct = PaliGemmaForConditionalGeneration(PaliGemmaConfig(**copy.deepcopy(TINY_CONFIG)), device="cuda", dtype=torch.bfloat16).eval()
    direct.load_state_dict(sd, strict=True)
    direct.tie_weights()
    inp = make_inputs(TINY_CONFIG, batch=2, prompt_len=5, seed=3)
    ref_t, ref_l = O.generate(sd, TINY_CONFIG, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], 5, return_logits=True)
    args = (inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), 5)
    loaded.language_model.deterministic_decode = direct.language_model.deterministic_decode = True
    _, a = loaded.generate(*args, return_logits=True, forced_tokens=ref_t)
    _, b = direct.generate(*args, return_logits=True, forced_tokens=ref_t)
    # (not bitwise: the few-token vision-tower GEMMs split K over CTAs and red.add their partials in arrival order)
    assert (a - b).abs().max().item() <= 1e-2 * b.abs().max().item(), "loader path and direct state-dict path hold different weights"
    for k, v in direct.state_dict().items():
        assert torch.equal(loaded.state_dict()[k], v), k
    err, absmax = (a.cpu() - ref_l).abs().max().item(), ref_l.abs().max().item()
    cos = torch.nn.functional.cosine_similarity(a.cpu().flatten().double(), ref_l.flatten().double(), dim=0).item()
    assert cos >= 0.9999 and err <= 0.01 * absmax, (err, absmax, cos)
    assert loaded.generate(*args).cpu().tolist() == ref_t.tolist()


@pytest.mark.gpu
def test_projector_bias_of_real_checkpoints_is_applied(tmp_path):
    """A checkpoint with multi_modal_projector.linear.bias: kept (with a warning), not silently dropped."""
    from paligemma_multimodal_system_b200.utils import load_hf_model
    sd = make_state_dict(TINY_CONFIG, "R1", seed=23)
    bias = torch.randn(TINY_CONFIG["projection_dim"]).bfloat16().float() * 0.5
    _write_checkpoint(str(tmp_path), sd, proj_bias=bias)
    with pytest.warns(UserWarning, match="projector"):
        model, _ = load_hf_model(str(tmp_path), device="cuda", tokenizer_loader=lambda p: _Tok())
    px = make_inputs(TINY_CONFIG, batch=1, seed=4)["pixel_values"].cuda()
    with_b = model.image_features(px).clone()
    model.set_projector_bias(None)
    without = model.image_features(px)
    # run-to-run noise of the split-K vision GEMMs is ~1e-3 of the feature scale; a dropped bias would be off by ~0.5
    assert (with_b - without - bias.cuda()).abs().max().item() <= 1e-2 * max(1.0, without.abs().max().item())
