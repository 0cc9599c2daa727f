"""SigLIP vision tower with the reference's module API (modeling_siglip.py:10-334) on sm_100a kernels.

The nn.Module tree below exists for API parity: constructor configs, parameter names / state-dict keys
(`vision_tower.model.encoder.layers.{i}.self_attn.key_proj.weight`, `...embeddings.positional_embeddings.weight`, ...),
`.to()`, `.eval()`, `load_state_dict`.  The compute path never calls the nn modules: `SiglipVisionModel.forward`
packs the parameters once (bf16, fused QKV) and runs

    im2col -> tcgen05 GEMM(+bias) -> +pos | per layer: LN -> fused QKV GEMM -> flash attention -> out_proj GEMM(+bias
    +residual) -> LN -> fc1 GEMM(+bias, gelu-tanh) -> fc2 GEMM(+bias +residual) | post-LN

through the C ABI (include/paligemma_b200.h).  The residual stream stays fp32 as in the reference.
"""
import torch
import torch.nn as nn

from . import _lib


class SiglipVisionConfig:
    """Same keyword arguments and defaults as modeling_siglip.py:10-38."""

    def __init__(self, image_size: int = 224, patch_size: int = 16, num_channels: int = 3, hidden_size: int = 768,
                 intermediate_size: int = 3072, num_hidden_layers: int = 12, num_attention_heads: int = 12,
                 attention_dropout: float = 0.0, layer_norm_eps: float = 1e-6, num_image_tokens: int = None, **kwargs):
        self.image_size = image_size
        self.patch_size = patch_size
        self.num_channels = num_channels
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.attention_dropout = attention_dropout
        self.layer_norm_eps = layer_norm_eps
        self.num_image_tokens = num_image_tokens


class _ParamsOnly(nn.Module):
    """Parameter container; compute happens in the packed kernels of the owning model."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: call the owning model's forward (sm_100a kernels), not this sub-module")


class SiglipAttention(_ParamsOnly):
    def __init__(self, config, **fk):
        super().__init__()
        D = config.hidden_size
        self.key_proj = nn.Linear(D, D, **fk)
        self.value_proj = nn.Linear(D, D, **fk)
        self.query_proj = nn.Linear(D, D, **fk)
        self.out_proj = nn.Linear(D, D, **fk)


class SiglipMLP(_ParamsOnly):
    def __init__(self, config, **fk):
        super().__init__()
        self.fc1 = nn.Linear(config.hidden_size, config.intermediate_size, **fk)
        self.fc2 = nn.Linear(config.intermediate_size, config.hidden_size, **fk)


class SiglipEncoderLayer(_ParamsOnly):
    def __init__(self, config, **fk):
        super().__init__()
        self.layer_norm1 = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps, **fk)
        self.self_attn = SiglipAttention(config, **fk)
        self.mlp = SiglipMLP(config, **fk)
        self.layer_norm2 = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps, **fk)


class SiglipEncoder(_ParamsOnly):
    def __init__(self, config, **fk):
        super().__init__()
        self.layers = nn.ModuleList([SiglipEncoderLayer(config, **fk) for _ in range(config.num_hidden_layers)])


class SiglipVisionEmbeddings(_ParamsOnly):
    def __init__(self, config, **fk):
        super().__init__()
        self.num_patches = (config.image_size // config.patch_size) ** 2
        self.patch_embedding = nn.Conv2d(config.num_channels, config.hidden_size, kernel_size=config.patch_size,
                                         stride=config.patch_size, padding="valid", **fk)
        self.positional_embeddings = nn.Embedding(self.num_patches, config.hidden_size, **fk)


class SiglipTransformer(_ParamsOnly):
    def __init__(self, config, **fk):
        super().__init__()
        self.config = config
        self.embeddings = SiglipVisionEmbeddings(config, **fk)
        self.encoder = SiglipEncoder(config, **fk)
        self.post_layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps, **fk)


def _bf16(t):
    return t.detach().to(device="cuda", dtype=torch.bfloat16).contiguous()


def _f32(t):
    return t.detach().to(device="cuda", dtype=torch.float32).contiguous()


class SiglipVisionModel(nn.Module):
    """forward(pixel_values [B,C,H,W]) -> [B, N, hidden] fp32 (modeling_siglip.py:324-334)."""

    def __init__(self, config: SiglipVisionConfig, device=None, dtype=None):
        super().__init__()
        self.config = config
        fk = {k: v for k, v in dict(device=device, dtype=dtype).items() if v is not None}
        self.model = SiglipTransformer(config, **fk)
        self._packed = None

    # -- packing -------------------------------------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def pack(self):
        """bf16 GEMM operands (QKV fused, conv flattened + K padded to a multiple of 64), fp32 biases / LN affine."""
        c = self.config
        m = self.model
        K = c.num_channels * c.patch_size ** 2
        Kp = (K + 63) // 64 * 64
        pw = torch.zeros(c.hidden_size, Kp, device="cuda", dtype=torch.bfloat16)
        pw[:, :K] = _bf16(m.embeddings.patch_embedding.weight).view(c.hidden_size, K)
        pk = dict(Kp=Kp, patch_w=pw, patch_b=_f32(m.embeddings.patch_embedding.bias),
                  pos=_f32(m.embeddings.positional_embeddings.weight), layers=[],
                  post_w=_f32(m.post_layernorm.weight), post_b=_f32(m.post_layernorm.bias))
        for l in m.encoder.layers:
            a = l.self_attn
            pk["layers"].append(dict(
                ln1_w=_f32(l.layer_norm1.weight), ln1_b=_f32(l.layer_norm1.bias),
                qkv_w=torch.cat([_bf16(a.query_proj.weight), _bf16(a.key_proj.weight), _bf16(a.value_proj.weight)], 0).contiguous(),
                qkv_b=torch.cat([_f32(a.query_proj.bias), _f32(a.key_proj.bias), _f32(a.value_proj.bias)], 0).contiguous(),
                out_w=_bf16(a.out_proj.weight), out_b=_f32(a.out_proj.bias),
                ln2_w=_f32(l.layer_norm2.weight), ln2_b=_f32(l.layer_norm2.bias),
                fc1_w=_bf16(l.mlp.fc1.weight), fc1_b=_f32(l.mlp.fc1.bias),
                fc2_w=_bf16(l.mlp.fc2.weight), fc2_b=_f32(l.mlp.fc2.bias)))
        self._packed = pk
        return pk

    # -- compute -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_features(self, pixel_values, out_bf16=False):
        """Returns the post-LN features as a flat [B*N, hidden] tensor (fp32, or bf16 for the projector GEMM)."""
        _lib.require_device()
        c = self.config
        pk = self._packed or self.pack()
        L = _lib.lib()
        st = _lib.stream()
        px = pixel_values.to(device="cuda", dtype=torch.float32).contiguous()
        B, C, H, W = px.shape
        if C != c.num_channels or H != c.image_size or W != c.image_size:
            raise ValueError(f"pixel_values must be [B,{c.num_channels},{c.image_size},{c.image_size}], got {tuple(px.shape)}")
        P, Dv, Fv, Hh = c.patch_size, c.hidden_size, c.intermediate_size, c.num_attention_heads
        dh = Dv // Hh
        N = (H // P) * (W // P)
        M = B * N
        dev = px.device
        patches = torch.empty(M, pk["Kp"], device=dev, dtype=torch.bfloat16)
        _lib.check(L.pg_im2col(px.data_ptr(), patches.data_ptr(), B, C, H, W, P, pk["Kp"], st), "pg_im2col")
        x = torch.empty(M, Dv, device=dev, dtype=torch.float32)
        # conv-as-GEMM with bias and the position embeddings (row n of the table for token b*N + n) in one epilogue
        _lib.gemm(patches, pk["patch_w"], x, mode=_lib.EPI_F32, bias=pk["patch_b"], resid=pk["pos"], resid_row_mod=N, swap=0)
        h = torch.empty(M, Dv, device=dev, dtype=torch.bfloat16)
        qkv = torch.empty(M, 3 * Dv, device=dev, dtype=torch.bfloat16)
        att = torch.empty(M, Dv, device=dev, dtype=torch.bfloat16)
        mid = torch.empty(M, Fv, device=dev, dtype=torch.bfloat16)
        scale = 1.0 / (dh ** 0.5)
        for lw in pk["layers"]:
            _lib.layernorm(x, lw["ln1_w"], lw["ln1_b"], c.layer_norm_eps, out_bf16=h)
            _lib.gemm(h, lw["qkv_w"], qkv, mode=_lib.EPI_BF16, bias=lw["qkv_b"], swap=0)
            _lib.check(L.pg_attention_prefill(
                qkv.data_ptr(), qkv.data_ptr() + 2 * Dv, qkv.data_ptr() + 4 * Dv, att.data_ptr(), B, Hh, N, N, dh, 1,
                N * 3 * Dv, 3 * Dv, 0, dh, N * 3 * Dv, 3 * Dv, dh, N * Dv, Dv, 0, dh, scale, st), "pg_attention_prefill")
            _lib.gemm_residual(att, lw["out_w"], x, bias=lw["out_b"])
            _lib.layernorm(x, lw["ln2_w"], lw["ln2_b"], c.layer_norm_eps, out_bf16=h)
            _lib.gemm(h, lw["fc1_w"], mid, mode=_lib.EPI_BF16, bias=lw["fc1_b"], act_gelu=True, swap=0)
            _lib.gemm_residual(mid, lw["fc2_w"], x, bias=lw["fc2_b"])
        if out_bf16:
            _lib.layernorm(x, pk["post_w"], pk["post_b"], c.layer_norm_eps, out_bf16=h)
            return h
        out = torch.empty_like(x)
        _lib.layernorm(x, pk["post_w"], pk["post_b"], c.layer_norm_eps, out_f32=out)
        return out

    def forward(self, x):
        B = x.shape[0]
        return self.forward_features(x).view(B, -1, self.config.hidden_size)
