"""load_hf_model with the reference's signature (utils.py:9-37): safetensors shards + config.json -> packed bf16 model.

SURVEY.md 8(f) "next" row 2.  Like the reference it loads with strict=False and ties the weights; unlike the reference
it also maps the Hugging Face SigLIP / projector key names onto the reference's parameter names (the reference silently
leaves the vision tower random because its names differ: `key_proj` vs `k_proj`, `positional_embeddings` vs
`position_embedding`, `vision_tower.model` vs `vision_tower.vision_model`)."""
import glob
import json
import os
from typing import Tuple

import torch

from .modeling_paligemma import PaliGemmaConfig, PaliGemmaForConditionalGeneration

_HF_TO_REF = (
    ("vision_tower.vision_model.", "vision_tower.model."),
    (".embeddings.position_embedding.", ".embeddings.positional_embeddings."),
    (".self_attn.k_proj.", ".self_attn.key_proj."),
    (".self_attn.v_proj.", ".self_attn.value_proj."),
    (".self_attn.q_proj.", ".self_attn.query_proj."),
)


def remap_hf_key(key: str) -> str:
    """HF `PaliGemmaForConditionalGeneration` checkpoint key -> reference module tree key (vision tower only: the
    language-model keys already agree)."""
    if key.startswith("vision_tower."):
        for a, b in _HF_TO_REF:
            key = key.replace(a, b)
    return key


def load_hf_model(model_path: str, device: str = "cuda") -> Tuple[PaliGemmaForConditionalGeneration, object]:
    from safetensors import safe_open
    from transformers import AutoTokenizer
    tokenizer = AutoTokenizer.from_pretrained(model_path, padding_side="right")
    assert tokenizer.padding_side == "right"
    tensors = {}
    for f in sorted(glob.glob(os.path.join(model_path, "*.safetensors"))):
        with safe_open(f, framework="pt", device="cpu") as sf:
            for key in sf.keys():
                tensors[remap_hf_key(key)] = sf.get_tensor(key)
    with open(os.path.join(model_path, "config.json")) as fh:
        config = PaliGemmaConfig(**json.load(fh))
    model = PaliGemmaForConditionalGeneration(config, device=device, dtype=torch.bfloat16)
    model.load_state_dict(tensors, strict=False)
    model.tie_weights()
    return model, tokenizer
