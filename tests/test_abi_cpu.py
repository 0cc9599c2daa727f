"""CPU-side checks: the C-ABI library loads and exports every symbol include/paligemma_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "paligemma_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from paligemma_multimodal_system_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/paligemma_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(syms), "ctypes SIGNATURES out of sync with the header"
    assert lib.pg_abi_version() >= 1


def test_product_fails_loudly_without_gpu_or_library(monkeypatch):
    import torch
    from paligemma_multimodal_system_b200 import _lib
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            _lib.require_device()
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpaligemma_b200.so")
    with pytest.raises(RuntimeError, match="missing"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "paligemma_multimodal_system_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{f} imports the oracle"
            assert "paligemma_oracle" not in src, f"{f} references the oracle module"


def test_config_classes_match_reference_defaults():
    from paligemma_multimodal_system_b200.modeling_gemma import GemmaConfig
    from paligemma_multimodal_system_b200.modeling_paligemma import PaliGemmaConfig
    from paligemma_multimodal_system_b200.modeling_siglip import SiglipVisionConfig
    from paligemma_multimodal_system_b200.random_init import TINY_CONFIG
    import copy
    v = SiglipVisionConfig()
    assert (v.image_size, v.patch_size, v.hidden_size, v.intermediate_size, v.num_hidden_layers, v.num_attention_heads,
            v.layer_norm_eps) == (224, 16, 768, 3072, 12, 12, 1e-6)  # modeling_siglip.py:11-21
    g = GemmaConfig()
    assert (g.rope_theta, g.max_position_encodings, g.head_dim, g.attention_bias) == (10000.0, 8192, 256, False)
    c = PaliGemmaConfig(**copy.deepcopy(TINY_CONFIG), some_unknown_key=1)  # unknown keys are swallowed (**kwargs)
    assert c.text_config.num_image_tokens == 256 and c.vision_config.projection_dim == 256
    assert c.vocab_size == 1281 and c.text_config.pad_token_id == 0 and c.ignore_index == -100
