"""Host-side pre-processing parity with the reference's documented behaviour (processing_paligemma.py:13-89,216-227)."""
import numpy as np
import pytest


class _Tok:
    bos_token = "<bos>"

    def __init__(self):
        self.vocab = {}

    def add_special_tokens(self, d):
        for t in d["additional_special_tokens"]:
            self.vocab.setdefault(t, len(self.vocab))

    def add_tokens(self, toks):
        for t in toks:
            self.vocab.setdefault(t, len(self.vocab))

    def convert_tokens_to_ids(self, t):
        return self.vocab[t]

    def __call__(self, strings, return_tensors=None, truncation=True, padding="longest"):
        import torch
        self.last = strings
        n = max(len(s) for s in strings)
        return {"input_ids": torch.zeros(len(strings), n, dtype=torch.long), "attention_mask": torch.ones(len(strings), n, dtype=torch.long)}


def test_process_images_shape_and_range():
    """The reference's (commented-out) smoke test: random 256x256 uint8 image -> (3, 224, 224), values in [-1, 1]."""
    from PIL import Image
    from paligemma_multimodal_system_b200.processing_paligemma import process_images
    img = Image.fromarray(np.random.randint(0, 256, (256, 256, 3), dtype=np.uint8))
    out = process_images([img], 224, 1 / 255.0, resampling=Image.Resampling.BICUBIC)
    assert out[0].shape == (3, 224, 224) and out[0].dtype == np.float32
    assert out[0].min() >= -1.0 and out[0].max() <= 1.0


def test_gemma_string_and_processor():
    from PIL import Image
    from paligemma_multimodal_system_b200.processing_paligemma import PaliGemmaProcessor, create_gemma_string
    assert create_gemma_string("caption en", 3, "<image>", "<bos>") == "<image><image><image><bos>caption en\n"
    tok = _Tok()
    proc = PaliGemmaProcessor(tok, 256, 224)
    assert tok.image_token_id == 0 and len(tok.vocab) == 1 + 128 + 1024 and tok.add_bos_token is False
    imgs = [Image.fromarray(np.zeros((300, 200, 3), dtype=np.uint8))] * 2
    out = proc(images=imgs, text=["caption en", "detect cat"])
    assert out["pixel_values"].shape == (2, 3, 224, 224)
    assert tok.last[1] == "<image>" * 256 + "<bos>detect cat\n"
    with pytest.raises(AssertionError):
        proc(images=imgs, text=["one prompt"])


def test_hf_key_remap():
    from paligemma_multimodal_system_b200.utils import remap_hf_key
    assert remap_hf_key("vision_tower.vision_model.encoder.layers.3.self_attn.k_proj.weight") == \
        "vision_tower.model.encoder.layers.3.self_attn.key_proj.weight"
    assert remap_hf_key("vision_tower.vision_model.embeddings.position_embedding.weight") == \
        "vision_tower.model.embeddings.positional_embeddings.weight"
    assert remap_hf_key("language_model.model.layers.0.self_attn.k_proj.weight") == "language_model.model.layers.0.self_attn.k_proj.weight"
