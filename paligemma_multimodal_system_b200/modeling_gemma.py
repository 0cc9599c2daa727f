"""Gemma decoder + KVCache with the reference's module API (modeling_gemma.py:8-533) on sm_100a kernels.

Compute path per layer (all through the C ABI of include/paligemma_b200.h):
  prefill : RMSNorm -> fused QKV GEMM (tcgen05) -> RoPE + paged-KV append -> flash attention (full, non-causal: the
            reference mask is all zeros, modeling_paligemma.py:154-156) -> o_proj GEMM (+fp32 residual) -> RMSNorm ->
            gate||up GEMM with gelu_tanh(g)*u epilogue -> down GEMM (+fp32 residual)
  decode  : same chain with weight-streaming (swap-AB) tcgen05 GEMMs, split-K fp32 reductions straight into the fp32
            residual stream, and split-KV attention over the paged bf16 cache.
The residual stream is fp32 (as in the reference); GEMM operands are bf16.
"""
import ctypes
import math
import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from .modeling_siglip import _ParamsOnly, _bf16, _f32

PAGE = 64  # tokens per KV page (= the decode attention key tile)


class KVCache:
    """Paged bf16 KV cache with the reference's interface (modeling_gemma.py:8-64): `update`, `num_items`, `k_cache`,
    `v_cache`.  Storage: k_pages / v_pages [layers, num_pages, 64, Hkv*dh] bf16 + page_table [B, max_pages] int32 + the
    per-row device counters the decode kernels read (position id, write slot, kv length)."""

    def __init__(self, reserve_tokens: int = 256):
        self.reserve_tokens = reserve_tokens  # head-room (in tokens) allocated beyond the prefill length
        self.k_pages = None
        self.v_pages = None
        self.page_table = None
        self.counters = None  # int32 [3, B]: pos, slot, kv_len  (device)
        self.image_feats = None  # projected image features of the request (decode-time <image> token parity)
        self._len = 0
        self._layer_len: List[int] = []
        self._geom = None  # (B, L, Hkv, dh)

    # -- reference API ---------------------------------------------------------------------------------------------
    def num_items(self) -> int:
        return self._len

    def update(self, key_states: torch.Tensor, value_states: torch.Tensor, layer_idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Appends [B, Hkv, s, dh] keys/values (already rotated) for `layer_idx`, returns the full K, V of that layer."""
        _lib.require_device()
        B, Hkv, s, dh = key_states.shape
        if self._geom is None:
            raise RuntimeError("KVCache.update before allocation: call allocate(B, layers, Hkv, dh, capacity) first")
        while layer_idx >= len(self._layer_len):
            self._layer_len.append(0)
        start = self._layer_len[layer_idx]
        self.ensure_capacity(start + s)
        W = Hkv * dh
        # [B,Hkv,s,dh] -> token-major [B*s, Hkv*dh]; the append kernel is pg_rope_kv_append with zero rotation
        k = key_states.to(device="cuda", dtype=torch.float32).permute(0, 2, 1, 3).reshape(B * s, W)
        v = value_states.to(device="cuda", dtype=torch.float32).permute(0, 2, 1, 3).reshape(B * s, W)
        qkv = torch.cat([k, v], 1).contiguous()  # "Hq = 0": only k and v heads
        pos = torch.zeros(B * s, device="cuda", dtype=torch.int32)
        inv = torch.zeros(dh // 2, device="cuda", dtype=torch.float32)
        base = torch.full((B,), start, device="cuda", dtype=torch.int32)
        dummy_q = torch.empty(1, device="cuda", dtype=torch.bfloat16)
        _lib.check(_lib.lib().pg_rope_kv_append(qkv.data_ptr(), 1, pos.data_ptr(), dummy_q.data_ptr(), 0, 0,
                                                self.k_pages[layer_idx].data_ptr(), self.v_pages[layer_idx].data_ptr(),
                                                self.page_table.data_ptr(), base.data_ptr(), B, s, 0, Hkv, dh, PAGE,
                                                self.page_table.shape[1], inv.data_ptr(), _lib.stream()), "pg_rope_kv_append")
        self._layer_len[layer_idx] = start + s
        self._len = self._layer_len[0]
        return self._dense(self.k_pages, layer_idx), self._dense(self.v_pages, layer_idx)

    @property
    def k_cache(self) -> List[torch.Tensor]:
        return [self._dense(self.k_pages, l) for l in range(len(self._layer_len))]

    @property
    def v_cache(self) -> List[torch.Tensor]:
        return [self._dense(self.v_pages, l) for l in range(len(self._layer_len))]

    # -- storage ---------------------------------------------------------------------------------------------------
    def allocate(self, B, layers, Hkv, dh, capacity_tokens):
        max_pages = (capacity_tokens + PAGE - 1) // PAGE
        self._geom = (B, layers, Hkv, dh)
        self.k_pages = torch.zeros(layers, B * max_pages, PAGE, Hkv * dh, device="cuda", dtype=torch.bfloat16)
        self.v_pages = torch.zeros_like(self.k_pages)
        self.page_table = torch.arange(B * max_pages, device="cuda", dtype=torch.int32).view(B, max_pages).contiguous()
        self.counters = torch.zeros(3, B, device="cuda", dtype=torch.int32)
        self._layer_len = []
        self._len = 0

    @property
    def capacity(self):
        return 0 if self.page_table is None else self.page_table.shape[1] * PAGE

    def ensure_capacity(self, tokens):
        if tokens <= self.capacity:
            return
        B, layers, Hkv, dh = self._geom
        old_k, old_v, old_pages = self.k_pages, self.v_pages, self.page_table.shape[1]
        new_pages = max((tokens + PAGE - 1) // PAGE, 2 * old_pages)
        self.k_pages = torch.zeros(layers, B * new_pages, PAGE, Hkv * dh, device="cuda", dtype=torch.bfloat16)
        self.v_pages = torch.zeros_like(self.k_pages)
        # pages of row b move to [b*new_pages, b*new_pages + old_pages)
        self.k_pages.view(layers, B, new_pages, PAGE, Hkv * dh)[:, :, :old_pages] = old_k.view(layers, B, old_pages, PAGE, Hkv * dh)
        self.v_pages.view(layers, B, new_pages, PAGE, Hkv * dh)[:, :, :old_pages] = old_v.view(layers, B, old_pages, PAGE, Hkv * dh)
        self.page_table = torch.arange(B * new_pages, device="cuda", dtype=torch.int32).view(B, new_pages).contiguous()

    def _dense(self, pages, layer):
        B, _, Hkv, dh = self._geom
        n = self._layer_len[layer] if layer < len(self._layer_len) else 0
        out = torch.empty(B, Hkv, n, dh, device="cuda", dtype=torch.bfloat16)
        if n > 0:
            _lib.check(_lib.lib().pg_kv_gather(pages[layer].data_ptr(), self.page_table.data_ptr(), out.data_ptr(), B, n, Hkv,
                                               dh, PAGE, self.page_table.shape[1], _lib.stream()), "pg_kv_gather")
        return out

    def _set_len(self, n, layers):
        self._len = n
        self._layer_len = [n] * layers


class GemmaConfig:
    """Same keyword arguments and defaults as modeling_gemma.py:68-99."""

    def __init__(self, rope_theta: float = 10000.0, max_position_encodings: int = 8192, rms_norm_eps: float = None,
                 hidden_size: int = None, num_hidden_layers: int = None, num_attention_heads: int = None,
                 num_key_value_heads: int = None, head_dim: int = 256, intermediate_size: int = None,
                 attention_bias: bool = False, attention_dropout: float = 0.0, pad_token_id: int = None,
                 vocab_size: int = None, **kwargs):
        self.rope_theta = rope_theta
        self.max_position_encodings = max_position_encodings
        self.rms_norm_eps = rms_norm_eps
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.num_key_value_heads = num_key_value_heads
        self.head_dim = head_dim
        self.intermediate_size = intermediate_size
        self.attention_bias = attention_bias
        self.attention_dropout = attention_dropout
        self.pad_token_id = pad_token_id
        self.vocab_size = vocab_size


class GemmaRMSNorm(_ParamsOnly):
    def __init__(self, dim: int, eps: float = 1e-6, **fk):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.zeros(dim, **fk))


class GemmaMLP(_ParamsOnly):
    def __init__(self, config, **fk):
        super().__init__()
        self.gate_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False, **fk)
        self.up_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False, **fk)
        self.down_proj = nn.Linear(config.intermediate_size, config.hidden_size, bias=False, **fk)


class GemmaAttention(_ParamsOnly):
    def __init__(self, config, layer_idx, **fk):
        super().__init__()
        assert config.num_attention_heads % config.num_key_value_heads == 0, \
            "number of Key/Value heads donot divide Number of Query Heads"
        if config.attention_bias:
            raise NotImplementedError("attention_bias=True is not on the PaliGemma path (modeling_gemma.py:80)")
        D, dh = config.hidden_size, config.head_dim
        self.layer_idx = layer_idx
        self.k_proj = nn.Linear(D, config.num_key_value_heads * dh, bias=False, **fk)
        self.v_proj = nn.Linear(D, config.num_key_value_heads * dh, bias=False, **fk)
        self.q_proj = nn.Linear(D, config.num_attention_heads * dh, bias=False, **fk)
        self.o_proj = nn.Linear(D, D, bias=False, **fk)  # requires Hq*dh == D (modeling_gemma.py:259)


class DecoderLayer(_ParamsOnly):
    def __init__(self, config, layer_idx=None, **fk):
        super().__init__()
        self.layer_idx = layer_idx
        self.input_layernorm = GemmaRMSNorm(config.hidden_size, **fk)
        self.self_attn = GemmaAttention(config, layer_idx, **fk)
        self.post_attention_layernorm = GemmaRMSNorm(config.hidden_size, **fk)
        self.mlp = GemmaMLP(config, **fk)


class GemmaModel(_ParamsOnly):
    def __init__(self, config, **fk):
        super().__init__()
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id, **fk)
        self.layers = nn.ModuleList([DecoderLayer(config, i, **fk) for i in range(config.num_hidden_layers)])
        self.norm = GemmaRMSNorm(config.hidden_size, **fk)


def _pick_cluster(tiles_m, total_kb, env=None, sms=148):
    """Cluster size (split-K ranks per output tile) for the decode GEMMs: enough CTAs to keep every SM streaming weights,
    a power of two <= 16, at least two k-blocks per rank."""
    if env and os.environ.get(env):
        return int(os.environ[env])
    s = 1
    while s < 16 and tiles_m * s < sms - 20 and total_kb // (2 * s) >= 2:
        s *= 2
    return s


def _pick_split(tiles_mn, total_kb, sms=148):
    """split-K so that roughly one wave of CTAs streams the weight matrix (decode GEMMs with few output tiles)."""
    if tiles_mn >= sms:
        return 1
    return max(1, min(total_kb, sms // tiles_mn))


class GemmaForCausalLM(nn.Module):
    """Parameter tree + packed-kernel compute of modeling_gemma.py:474-533."""

    def __init__(self, config: GemmaConfig, device=None, dtype=None):
        super().__init__()
        fk = {k: v for k, v in dict(device=device, dtype=dtype).items() if v is not None}
        self.text_config = config
        self.vocab_size = config.vocab_size
        self.hidden_size = config.hidden_size
        if config.num_attention_heads * config.head_dim != config.hidden_size:
            raise ValueError("o_proj is Linear(hidden, hidden): num_attention_heads*head_dim must equal hidden_size")
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, **fk)  # biased, as in modeling_gemma.py:484
        self.model = GemmaModel(config, **fk)
        self._packed = None
        self._ws = {}
        self._step_maps = {}
        self._barrier_state = None

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def tie_weights(self):
        self.lm_head.weight = self.model.embed_tokens.weight
        self._packed = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    # -- packing ---------------------------------------------------------------------------------------------------
    def pack(self):
        c = self.text_config
        L = _lib.lib()
        embed = _bf16(self.model.embed_tokens.weight)
        tied = self.lm_head.weight is self.model.embed_tokens.weight
        inv_freq = 1.0 / (c.rope_theta ** (torch.arange(0, c.head_dim, 2, dtype=torch.int64).float() / c.head_dim))
        pk = dict(embed=embed, head_w=embed if tied else _bf16(self.lm_head.weight), head_b=_f32(self.lm_head.bias),
                  norm_w=_f32(self.model.norm.weight), inv_freq=inv_freq.to("cuda"), layers=[])
        F, D = c.intermediate_size, c.hidden_size
        if F % 64 != 0:
            raise ValueError("intermediate_size must be a multiple of 64 for the packed gate||up layout")
        for l in self.model.layers:
            a = l.self_attn
            gate, up = _bf16(l.mlp.gate_proj.weight), _bf16(l.mlp.up_proj.weight)
            gu = torch.empty(2 * F, D, device="cuda", dtype=torch.bfloat16)
            _lib.check(L.pg_pack_gate_up(gate.data_ptr(), up.data_ptr(), gu.data_ptr(), F, D, _lib.stream()), "pg_pack_gate_up")
            pk["layers"].append(dict(
                ln1=_f32(l.input_layernorm.weight), ln2=_f32(l.post_attention_layernorm.weight),
                qkv_w=torch.cat([_bf16(a.q_proj.weight), _bf16(a.k_proj.weight), _bf16(a.v_proj.weight)], 0).contiguous(),
                o_w=_bf16(a.o_proj.weight), gu_w=gu, down_w=_bf16(l.mlp.down_proj.weight)))
            del gate, up
        pk["ln1_all"] = torch.stack([lw["ln1"] for lw in pk["layers"]]).contiguous()
        pk["ln2_all"] = torch.stack([lw["ln2"] for lw in pk["layers"]]).contiguous()
        torch.cuda.synchronize()
        self._packed = pk
        self._step_maps = {}
        return pk

    def _buf(self, name, shape, dtype):
        t = self._ws.get(name)
        if t is None or t.shape != tuple(shape) or t.dtype != dtype:
            t = torch.empty(*shape, device="cuda", dtype=dtype)
            self._ws[name] = t
        return t

    # -- prefill ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prefill(self, h, pos, B, S, kv_cache: Optional[KVCache], last_only: bool, reserve_tokens: int = 0, lens=None,
                page_rows=None):
        """h fp32 [B*S, D] (merged, scaled embeddings; overwritten), pos int32 [B*S] -> logits fp32 [B, S|1, V].

        Ragged batches (serving.py): `lens` int32 [B] on the device = true prompt length of each row (rows are right-padded
        to S; row b only attends to its first lens[b] keys and `last_only` picks position lens[b]-1); `page_rows` int32
        [B, max_pages] = page-table rows (into the pages of an already allocated `kv_cache`) that receive the keys/values of
        the B rows; the cache's own page table and length bookkeeping are left to the caller."""
        c = self.text_config
        pk = self._packed or self.pack()
        L, st = _lib.lib(), _lib.stream()
        D, F, Hq, Hkv, dh, V = c.hidden_size, c.intermediate_size, c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.vocab_size
        T = B * S
        dev = h.device
        have_cache = kv_cache is not None
        page_table = None
        if have_cache and page_rows is not None:
            if kv_cache._geom is None or page_rows.shape[1] * PAGE < S or page_rows.shape[0] != B:
                raise ValueError("prefill through page_rows needs an allocated KVCache and rows that hold the padded prompt")
            page_table = page_rows.to(torch.int32).contiguous()
            slot_base = torch.zeros(B, device=dev, dtype=torch.int32)
        elif have_cache:
            if kv_cache._geom is None or kv_cache._geom != (B, c.num_hidden_layers, Hkv, dh):
                kv_cache.allocate(B, c.num_hidden_layers, Hkv, dh, S + max(kv_cache.reserve_tokens, reserve_tokens))
            kv_cache.ensure_capacity(S + 1)
            slot_base = torch.zeros(B, device=dev, dtype=torch.int32)
            page_table = kv_cache.page_table
        hn = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
        qkv = torch.empty(T, (Hq + 2 * Hkv) * dh, device=dev, dtype=torch.bfloat16)
        q = torch.empty(T, Hq * dh, device=dev, dtype=torch.bfloat16)
        k = torch.empty(T, Hkv * dh, device=dev, dtype=torch.bfloat16)
        v = torch.empty(T, Hkv * dh, device=dev, dtype=torch.bfloat16)
        att = torch.empty(T, Hq * dh, device=dev, dtype=torch.bfloat16)
        mid = torch.empty(T, F, device=dev, dtype=torch.bfloat16)
        G = Hq // Hkv
        scale = 1.0 / math.sqrt(dh)
        for li, lw in enumerate(pk["layers"]):
            _lib.rmsnorm(h, lw["ln1"], hn)
            _lib.gemm(hn, lw["qkv_w"], qkv, mode=_lib.EPI_BF16, swap=0 if T > 128 else 1)
            _lib.check(L.pg_rope_kv_append(
                qkv.data_ptr(), 0, pos.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(),
                kv_cache.k_pages[li].data_ptr() if have_cache else 0, kv_cache.v_pages[li].data_ptr() if have_cache else 0,
                page_table.data_ptr() if have_cache else 0, slot_base.data_ptr() if have_cache else 0,
                B, S, Hq, Hkv, dh, PAGE, page_table.shape[1] if have_cache else 0, pk["inv_freq"].data_ptr(), st),
                "pg_rope_kv_append")
            # MQA/GQA: the G query heads of a KV head are consecutive rows of one attention problem (no repeat_kv)
            if lens is None:
                _lib.check(L.pg_attention_prefill(
                    q.data_ptr(), k.data_ptr(), v.data_ptr(), att.data_ptr(), B, Hkv, S * G, S, dh, G,
                    S * Hq * dh, Hq * dh, dh, G * dh, S * Hkv * dh, Hkv * dh, dh,
                    S * Hq * dh, Hq * dh, dh, G * dh, scale, st), "pg_attention_prefill")
            else:
                _lib.check(L.pg_attention_prefill_varlen(
                    q.data_ptr(), k.data_ptr(), v.data_ptr(), att.data_ptr(), lens.data_ptr(), B, Hkv, S * G, S, dh, G,
                    S * Hq * dh, Hq * dh, dh, G * dh, S * Hkv * dh, Hkv * dh, dh,
                    S * Hq * dh, Hq * dh, dh, G * dh, scale, st), "pg_attention_prefill_varlen")
            _lib.gemm_residual(att, lw["o_w"], h)
            _lib.rmsnorm(h, lw["ln2"], hn)
            _lib.gemm(hn, lw["gu_w"], mid, mode=_lib.EPI_GEGLU, swap=0 if T > 128 else 1)
            _lib.gemm_residual(mid, lw["down_w"], h)
        if have_cache and page_rows is None:
            kv_cache._set_len(S, c.num_hidden_layers)
        if last_only:
            if lens is None:
                last = h.view(B, S, D)[:, -1, :].contiguous()
            else:
                last = h.view(B, S, D)[torch.arange(B, device=dev), lens.long() - 1].contiguous()
            ln = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
            _lib.rmsnorm(last, pk["norm_w"], ln)
            logits = torch.empty(B, V, device=dev, dtype=torch.float32)
            _lib.gemm(ln, pk["head_w"], logits, mode=_lib.EPI_F32, bias=pk["head_b"], swap=1 if B <= 128 else 0)
            return logits.view(B, 1, V)
        _lib.rmsnorm(h, pk["norm_w"], hn)
        logits = torch.empty(T, V, device=dev, dtype=torch.float32)
        _lib.gemm(hn, pk["head_w"], logits, mode=_lib.EPI_F32, bias=pk["head_b"], swap=0 if T > 128 else 1)
        return logits.view(B, S, V)

    # -- decode ----------------------------------------------------------------------------------------------------
    def decode_buffers(self, B, private: bool = False):
        """Activation buffers of one decode step.  Cached per batch size: captured CUDA graphs (generate(), serving.py) hold
        these addresses, so a buffer must never be replaced once handed out; `private` returns a fresh set owned by the
        caller."""
        c = self.text_config
        D, F, Hq, Hkv, dh, V = c.hidden_size, c.intermediate_size, c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.vocab_size
        shapes = dict(h=((B, D), torch.float32), hn=((B, D), torch.bfloat16), hb=((B, D), torch.bfloat16),
                      ss=((2 * c.num_hidden_layers + 1, B), torch.float32), qkv=((B, (Hq + 2 * Hkv) * dh), torch.float32),
                      att=((B, Hq * dh), torch.bfloat16), mid=((B, F), torch.bfloat16), logits=((B, V), torch.float32))
        if private:
            return {k: torch.empty(*shp, device="cuda", dtype=dt) for k, (shp, dt) in shapes.items()}
        return {k: self._buf(f"d_{k}_{B}", shp, dt) for k, (shp, dt) in shapes.items()}

    @torch.no_grad()
    def decode_prologue(self, bufs, B, tokens_i32=None, img=None, img_scale=1.0, pad_token=-1, image_token=-1):
        """Token embedding (or the fp32 rows already in bufs['h'] when tokens_i32 is None) -> operands of the first fused
        RMSNorm (hb = bf16(h * (1 + ln1_0)), ss[0] = sum h^2); zeroes the remaining sum-of-squares accumulators."""
        c = self.text_config
        pk = self._packed or self.pack()
        D = c.hidden_size
        ss = bufs["ss"]
        _lib.check(_lib.lib().pg_decode_prologue(
            _lib.ptr(tokens_i32), pk["embed"].data_ptr(), _lib.ptr(img), bufs["h"].data_ptr(), bufs["hb"].data_ptr(),
            ss[0].data_ptr(), pk["layers"][0]["ln1"].data_ptr(), ss[1].data_ptr(), ss.numel() - ss.shape[1], B, D,
            0 if img is None else img.shape[1], D ** 0.5, img_scale, pad_token if pad_token is not None else -1, image_token,
            _lib.stream()), "pg_decode_prologue")

    @torch.no_grad()
    def decode_layers(self, bufs, kv_cache: KVCache, B):
        """One decode step over all layers; reads bufs['h'] (fp32 embeddings), leaves fp32 logits in bufs['logits'].
        Seven launches per layer: RMSNorm, QKV GEMM, fused RoPE + KV append + attention, O GEMM, RMSNorm, gate||up GEGLU
        GEMM, down GEMM; the three small-output GEMMs split K over CTAs and red.add their fp32 partials straight into the
        residual stream / the pre-zeroed qkv buffer.  Every launch reads its sizes from device counters, so the sequence
        can be captured in a CUDA graph.  (PG_DECODE_CLUSTER=1 selects the variant whose split-K reduction runs inside
        thread-block clusters and folds the RMSNorms into the GEMM epilogues: measured slower on B200)"""
        if os.environ.get("PG_DECODE_CLUSTER", "0") == "1":
            return self.decode_layers_cluster(bufs, kv_cache, B)
        c = self.text_config
        pk = self._packed or self.pack()
        L, st = _lib.lib(), _lib.stream()
        D, F, Hq, Hkv, dh, V = c.hidden_size, c.intermediate_size, c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.vocab_size
        h, hn, qkv, att, mid = bufs["h"], bufs["hn"], bufs["qkv"], bufs["att"], bufs["mid"]
        pos, kvl = kv_cache.counters[0], kv_cache.counters[2]
        max_pages = kv_cache.page_table.shape[1]
        scale = 1.0 / math.sqrt(dh)
        W = (Hq + 2 * Hkv) * dh
        sp_qkv = _pick_split((W + 127) // 128, D // 64)
        sp_o = _pick_split((D + 127) // 128, D // 64)
        sp_down = _pick_split((D + 127) // 128, F // 64, sms=296)
        if getattr(self, "deterministic_decode", False):
            # one CTA per output tile: the fp32 red.add has a single writer per element, so the step is bitwise
            # reproducible run to run (the split-K partials otherwise arrive in a different order every launch)
            sp_qkv = sp_o = sp_down = 1
        for li, lw in enumerate(pk["layers"]):
            _lib.rmsnorm(h, lw["ln1"], hn, zero_buf=qkv)
            _lib.gemm(hn, lw["qkv_w"], qkv, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sp_qkv)
            # RoPE + KV append + split-KV attention + combine: one launch
            _lib.check(L.pg_attention_decode_fused(
                qkv.data_ptr(), pos.data_ptr(), kvl.data_ptr(), pk["inv_freq"].data_ptr(), kv_cache.k_pages[li].data_ptr(),
                kv_cache.v_pages[li].data_ptr(), kv_cache.page_table.data_ptr(), att.data_ptr(), B, Hq, Hkv, dh, PAGE,
                kv_cache.k_pages.shape[1], max_pages, scale, st), "pg_attention_decode_fused")
            _lib.gemm(att, lw["o_w"], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sp_o)
            _lib.rmsnorm(h, lw["ln2"], hn)
            _lib.gemm(hn, lw["gu_w"], mid, mode=_lib.EPI_GEGLU, swap=1)
            _lib.gemm(mid, lw["down_w"], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sp_down)
        _lib.rmsnorm(h, pk["norm_w"], hn)
        _lib.gemm(hn, pk["head_w"], bufs["logits"], mode=_lib.EPI_F32, bias=pk["head_b"], swap=1)
        return bufs["logits"]

    @torch.no_grad()
    def decode_layers_cluster(self, bufs, kv_cache: KVCache, B):
        """Variant of decode_layers with five launches per layer: QKV GEMM (cluster split-K, input RMSNorm applied as a
        per-token factor), attention, O GEMM (+residual, emits the post-attention norm operands), gate||up GEGLU GEMM, down
        GEMM (+residual, emits the next layer's norm operands).  Needs decode_prologue() to have filled hb / ss[0]."""
        c = self.text_config
        pk = self._packed or self.pack()
        L, st = _lib.lib(), _lib.stream()
        D, F, Hq, Hkv, dh, V = c.hidden_size, c.intermediate_size, c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.vocab_size
        h, hb, ss, qkv, att, mid = bufs["h"], bufs["hb"], bufs["ss"], bufs["qkv"], bufs["att"], bufs["mid"]
        pos, kvl = kv_cache.counters[0], kv_cache.counters[2]
        max_pages = kv_cache.page_table.shape[1]
        scale = 1.0 / math.sqrt(dh)
        W = (Hq + 2 * Hkv) * dh
        eps = 1e-6
        s_qkv = _pick_cluster((W + 127) // 128, D // 64, "PG_S_QKV")
        s_o = _pick_cluster((D + 127) // 128, D // 64, "PG_S_O")
        s_down = _pick_cluster((D + 127) // 128, F // 64, "PG_S_DOWN")
        n_layers = len(pk["layers"])
        for li, lw in enumerate(pk["layers"]):
            _lib.gemm_decode(hb, lw["qkv_w"], qkv, mode=_lib.DEC_F32, cluster_k=s_qkv, ss_in=ss[2 * li], norm_dim=D, eps=eps)
            _lib.check(L.pg_attention_decode_fused(
                qkv.data_ptr(), pos.data_ptr(), kvl.data_ptr(), pk["inv_freq"].data_ptr(), kv_cache.k_pages[li].data_ptr(),
                kv_cache.v_pages[li].data_ptr(), kv_cache.page_table.data_ptr(), att.data_ptr(), B, Hq, Hkv, dh, PAGE,
                kv_cache.k_pages.shape[1], max_pages, scale, st), "pg_attention_decode_fused")
            _lib.gemm_decode(att, lw["o_w"], h, mode=_lib.DEC_RESID_NORM, cluster_k=s_o, hb=hb, norm_w=lw["ln2"], ss_out=ss[2 * li + 1])
            _lib.gemm_colnorm(hb, lw["gu_w"], mid, mode=_lib.EPI_GEGLU, ss_in=ss[2 * li + 1], norm_dim=D, eps=eps)
            next_w = pk["layers"][li + 1]["ln1"] if li + 1 < n_layers else pk["norm_w"]
            _lib.gemm_decode(mid, lw["down_w"], h, mode=_lib.DEC_RESID_NORM, cluster_k=s_down, hb=hb, norm_w=next_w, ss_out=ss[2 * li + 2])
        _lib.gemm_colnorm(hb, pk["head_w"], bufs["logits"], mode=_lib.EPI_F32, bias=pk["head_b"], ss_in=ss[2 * n_layers], norm_dim=D, eps=eps)
        return bufs["logits"]

    # -- one-kernel decode step ------------------------------------------------------------------------------------------
    def megakernel_ok(self, B):
        c = self.text_config
        # opt-in: measured slower than the PDL-chained per-op kernels (2.19 vs 1.74 ms / step at 3B, 64 sequences): a grid
        # barrier costs ~2.3 us, about the same as a programmatic-dependent-launch kernel boundary
        return (os.environ.get("PG_MEGAKERNEL", "0") == "1" and B * c.num_key_value_heads <= 148 and B <= 64
                and c.head_dim in (64, 256) and c.num_attention_heads // c.num_key_value_heads <= 8)

    @torch.no_grad()
    def decode_step(self, bufs, kv_cache: KVCache, B, tokens_i32, img, img_scale, pad_token, image_token):
        """Embeds `tokens_i32` [B] and runs every layer + final norm + lm_head; fp32 logits land in bufs['logits'].
        B <= 64: ONE persistent cooperative kernel (csrc/decode_step.cu); otherwise the per-op kernels."""
        c = self.text_config
        pk = self._packed or self.pack()
        L = _lib.lib()
        D, F, Hq, Hkv, dh, V = c.hidden_size, c.intermediate_size, c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.vocab_size
        if not self.megakernel_ok(B):
            if os.environ.get("PG_DECODE_CLUSTER", "0") == "1":
                self.decode_prologue(bufs, B, tokens_i32, img, img_scale, pad_token, image_token)
            else:
                _lib.check(L.pg_embed_tokens(tokens_i32.data_ptr(), pk["embed"].data_ptr(), _lib.ptr(img), bufs["h"].data_ptr(), B, D,
                                             0 if img is None else img.shape[1], D ** 0.5, img_scale, pad_token, image_token,
                                             _lib.stream()), "pg_embed_tokens")
            return self.decode_layers(bufs, kv_cache, B)
        key = (B, bufs["hn"].data_ptr(), bufs["att"].data_ptr(), bufs["mid"].data_ptr())
        maps = self._step_maps.get(key)
        if maps is None:
            n = 4 * c.num_hidden_layers + 4
            host = ctypes.create_string_buffer(n * 128)
            wl = []
            for lw in pk["layers"]:
                wl += [lw["qkv_w"].data_ptr(), lw["o_w"].data_ptr(), lw["gu_w"].data_ptr(), lw["down_w"].data_ptr()]
            wl.append(pk["head_w"].data_ptr())
            warr = (ctypes.c_void_p * len(wl))(*wl)
            _lib.check(L.pg_decode_step_encode_maps(ctypes.addressof(host), ctypes.addressof(warr), bufs["hn"].data_ptr(),
                                                    bufs["att"].data_ptr(), bufs["mid"].data_ptr(), c.num_hidden_layers, B, D, F,
                                                    Hq, Hkv, dh, V), "pg_decode_step_encode_maps")
            maps = torch.frombuffer(bytearray(host.raw), dtype=torch.uint8).cuda()
            self._step_maps = {key: maps}
        if self._barrier_state is None:
            self._barrier_state = torch.zeros(256, device="cuda", dtype=torch.int32)
        W = (Hq + 2 * Hkv) * dh
        a = _lib.DecodeStepArgs()
        a.tensor_maps = maps.data_ptr()
        a.L, a.B, a.D, a.F, a.Hq, a.Hkv, a.dh, a.V = c.num_hidden_layers, B, D, F, Hq, Hkv, dh, V
        a.split_qkv = _pick_split((W + 127) // 128, D // 64)
        a.split_o = _pick_split((D + 127) // 128, D // 64)
        a.split_down = _pick_split((D + 127) // 128, F // 64)
        a.cur_tok, a.embed, a.img = tokens_i32.data_ptr(), pk["embed"].data_ptr(), _lib.ptr(img)
        a.n_img = 0 if img is None else img.shape[1]
        a.text_scale, a.img_scale, a.pad_token, a.image_token = D ** 0.5, img_scale, pad_token, image_token
        a.h, a.hn, a.qkv, a.att, a.mid, a.logits = (bufs[k].data_ptr() for k in ("h", "hn", "qkv", "att", "mid", "logits"))
        a.ln1, a.ln2, a.norm_w, a.head_b, a.eps = pk["ln1_all"].data_ptr(), pk["ln2_all"].data_ptr(), pk["norm_w"].data_ptr(), pk["head_b"].data_ptr(), 1e-6
        a.k_pages, a.v_pages = kv_cache.k_pages.data_ptr(), kv_cache.v_pages.data_ptr()
        a.layer_stride = kv_cache.k_pages.stride(0)
        a.page_table, a.pos, a.kv_len = kv_cache.page_table.data_ptr(), kv_cache.counters[0].data_ptr(), kv_cache.counters[2].data_ptr()
        a.inv_freq, a.max_pages, a.page_size, a.scale = pk["inv_freq"].data_ptr(), kv_cache.page_table.shape[1], PAGE, 1.0 / math.sqrt(dh)
        a.barrier_state = self._barrier_state.data_ptr()
        a.trace = _lib.ptr(getattr(self, "_trace", None))
        a.trace_cta = getattr(self, "_trace_cta", 0)
        _lib.check(L.pg_decode_step(ctypes.byref(a), _lib.stream()), "pg_decode_step")
        return bufs["logits"]

    def forward(self, input_embeds=None, position_ids=None, attention_mask=None, kv_cache=None):
        """Reference signature (modeling_gemma.py:501-533): input_embeds [B,S,D] UNSCALED (the sqrt(D) normaliser is
        applied here, :510-511), position_ids [B,S]; the additive attention_mask is all zeros on this path and ignored."""
        _lib.require_device()
        B, S, D = input_embeds.shape
        h = (input_embeds.to(device="cuda", dtype=torch.float32) * torch.tensor(D ** 0.5, dtype=torch.float32)).reshape(B * S, D).contiguous()
        pos = position_ids.to(device="cuda", dtype=torch.int32).reshape(-1).contiguous()
        if kv_cache is not None and kv_cache.num_items() > 0:
            if S != 1:
                raise AssertionError("Generation Phase more than one token CAN'T be input")
            n = kv_cache.num_items()
            kv_cache.ensure_capacity(n + 1)
            kv_cache.counters[0].copy_(pos)
            kv_cache.counters[1].fill_(n)
            kv_cache.counters[2].fill_(n + 1)
            bufs = self.decode_buffers(B)
            bufs["h"].copy_(h)
            if os.environ.get("PG_DECODE_CLUSTER", "0") == "1":
                self.decode_prologue(bufs, B)
            logits = self.decode_layers(bufs, kv_cache, B).clone().view(B, 1, -1)
            kv_cache._set_len(n + 1, self.text_config.num_hidden_layers)
        else:
            logits = self.prefill(h, pos, B, S, kv_cache, last_only=False)
        out = {"logits": logits}
        if kv_cache is not None:
            out["kv_cache"] = kv_cache
        return out
