"""load_hf_model with the reference's signature (utils.py:9-37): safetensors shards + config.json -> packed bf16 model.

SURVEY.md 8(f) "next" row 2.  Like the reference it reads every `*.safetensors` shard, builds the model from
`config.json`, loads the tensors and ties the weights.  Where the reference's `strict=False` hides real differences, this
loader accounts for every key instead:

  * Hugging Face SigLIP / projector key names are mapped onto the reference's parameter names (the reference silently
    leaves its vision tower at random init because the names differ: `key_proj` vs `k_proj`, `positional_embeddings`
    vs `position_embedding`, `vision_tower.model` vs `vision_tower.vision_model`);
  * `language_model.lm_head.weight` may be absent (tied to the embedding, modeling_gemma.py:492-499);
  * `language_model.lm_head.bias` does not exist in a PaliGemma checkpoint: the reference keeps nn.Linear's unseeded random
    init there (+-0.022 of run-to-run noise on every logit, modeling_gemma.py:484); here it is ZEROED, which is what
    the checkpoint means, and makes greedy decoding reproducible;
  * `multi_modal_projector.linear.bias` exists in real checkpoints while the reference builds the projector with
    bias=False (modeling_paligemma.py:57) and drops it; here the bias is kept and applied in the projector GEMM
    epilogue, with a warning that the reference would have dropped it;
  * anything else missing or unexpected raises.
"""
import glob
import json
import os
import warnings
from typing import Callable, Dict, Optional, Tuple

import torch

from .modeling_paligemma import PaliGemmaConfig, PaliGemmaForConditionalGeneration

_HF_TO_REF = (
    ("vision_tower.vision_model.", "vision_tower.model."),
    (".embeddings.position_embedding.", ".embeddings.positional_embeddings."),
    (".self_attn.k_proj.", ".self_attn.key_proj."),
    (".self_attn.v_proj.", ".self_attn.value_proj."),
    (".self_attn.q_proj.", ".self_attn.query_proj."),
)

# checkpoint keys that have no parameter on this path (and none in the reference): SigLIP's position-id buffer of older
# transformers releases and its pooling head, which PaliGemma does not use
_IGNORED_SUFFIXES = (".embeddings.position_ids",)
_IGNORED_PREFIXES = ("vision_tower.model.head.",)
_MAY_BE_MISSING = ("language_model.lm_head.weight", "language_model.lm_head.bias")
PROJECTOR_BIAS = "multi_modal_projector.linear.bias"


def remap_hf_key(key: str) -> str:
    """HF `PaliGemmaForConditionalGeneration` checkpoint key -> reference module tree key (vision tower only: the
    language-model keys already agree)."""
    if key.startswith("vision_tower."):
        for a, b in _HF_TO_REF:
            key = key.replace(a, b)
    return key


def read_hf_checkpoint(model_path: str) -> Tuple[Dict[str, torch.Tensor], dict]:
    """Every tensor of every `*.safetensors` shard under reference-style keys (CPU), and the parsed config.json."""
    from safetensors import safe_open
    files = sorted(glob.glob(os.path.join(model_path, "*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {model_path}")
    tensors = {}
    for f in files:
        with safe_open(f, framework="pt", device="cpu") as sf:
            for key in sf.keys():
                tensors[remap_hf_key(key)] = sf.get_tensor(key)
    with open(os.path.join(model_path, "config.json")) as fh:
        return tensors, json.load(fh)


def load_hf_model(model_path: str, device: str = "cuda", tokenizer_loader: Optional[Callable] = None
                  ) -> Tuple[PaliGemmaForConditionalGeneration, object]:
    """(model, tokenizer) as utils.py:9-37.  `tokenizer_loader(model_path)` replaces AutoTokenizer.from_pretrained (tests,
    deployments that ship their own tokenizer object)."""
    if tokenizer_loader is None:
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(model_path, padding_side="right")
    else:
        tokenizer = tokenizer_loader(model_path)
    assert tokenizer.padding_side == "right"
    tensors, cfg = read_hf_checkpoint(model_path)
    tensors = {k: v for k, v in tensors.items() if not k.endswith(_IGNORED_SUFFIXES) and not k.startswith(_IGNORED_PREFIXES)}
    config = PaliGemmaConfig(**cfg)
    model = PaliGemmaForConditionalGeneration(config, device=device, dtype=torch.bfloat16).eval()
    proj_bias = tensors.pop(PROJECTOR_BIAS, None)
    missing, unexpected = model.load_state_dict(tensors, strict=False)
    missing = [k for k in missing if k not in _MAY_BE_MISSING]
    if missing or unexpected:
        raise KeyError(f"checkpoint does not match the PaliGemma module tree: missing {sorted(missing)[:8]}"
                       f"{' ...' if len(missing) > 8 else ''}, unexpected {sorted(unexpected)[:8]}{' ...' if len(unexpected) > 8 else ''}")
    with torch.no_grad():
        if "language_model.lm_head.bias" not in tensors:
            model.language_model.lm_head.bias.zero_()
        if proj_bias is not None:
            warnings.warn("checkpoint has multi_modal_projector.linear.bias: applied here; the reference builds the projector "
                          "with bias=False and drops it (modeling_paligemma.py:57)")
            model.set_projector_bias(proj_bias)
    model.tie_weights()
    return model, tokenizer
