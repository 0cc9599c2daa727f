"""Gemma decoder + KVCache with the reference's module API (modeling_gemma.py:8-533) on sm_100a kernels.

Compute path per layer (all through the C ABI of include/paligemma_b200.h):
  prefill : RMSNorm -> fused QKV GEMM (tcgen05) -> RoPE + paged-KV append -> flash attention (full, non-causal: the
            reference mask is all zeros, modeling_paligemma.py:154-156) -> o_proj GEMM (+fp32 residual) -> RMSNorm ->
            gate||up GEMM with gelu_tanh(g)*u epilogue -> down GEMM (+fp32 residual)
  decode  : same chain with weight-streaming (swap-AB) tcgen05 GEMMs, split-K fp32 reductions straight into the fp32
            residual stream, and split-KV attention over the paged bf16 cache.
The residual stream is fp32 (as in the reference); GEMM operands are bf16.
"""
import math
import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from .modeling_siglip import _ParamsOnly, _bf16, _f32

PAGE = 64  # tokens per KV page (= the decode attention key tile)
MAX_DECODE_BATCH = 128  # rows of one decode step (the batch sits on the UMMA N axis of the weight-streaming GEMMs)


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
        self.image_feats = None  # projected image features of the request (decode-time <image> token parity): [B, n >= 1, D]
        self.image_feats_scaled = False  # True: the rows already carry hidden_size**-0.5 * sqrt(hidden) (written by _embed_prompt)
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
        if self._geom is None:  # the reference's KVCache() is usable as constructed: its first update creates the layer
            self.allocate(B, layer_idx + 1, Hkv, dh, s + self.reserve_tokens)
        if self._geom[0] != B or self._geom[2:] != (Hkv, dh):
            raise ValueError(f"KVCache holds [B, Hkv, dh] = {(self._geom[0],) + self._geom[2:]}, update got {(B, Hkv, dh)}")
        if layer_idx >= self._geom[1]:
            self._grow_layers(layer_idx + 1)
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

    def _grow_layers(self, layers):
        B, old_layers, Hkv, dh = self._geom
        extra = torch.zeros(layers - old_layers, *self.k_pages.shape[1:], device="cuda", dtype=torch.bfloat16)
        self.k_pages = torch.cat([self.k_pages, extra], 0)
        self.v_pages = torch.cat([self.v_pages, torch.zeros_like(extra)], 0)
        self._geom = (B, layers, Hkv, dh)

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


def _pick_split(tiles_mn, total_kb, sms):
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
        self.fused_qkv_rope = None  # None = automatic (prefill()); True / False force the q/k/v-epilogue RoPE + KV append path

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
        torch.cuda.synchronize()
        self._packed = pk
        return pk

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
        W = (Hq + 2 * Hkv) * dh
        hn = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
        qkv = torch.empty(T, W, device=dev, dtype=torch.bfloat16)
        att = torch.empty(T, Hq * dh, device=dev, dtype=torch.bfloat16)
        mid = torch.empty(T, F, device=dev, dtype=torch.bfloat16)
        G = Hq // Hkv
        scale = 1.0 / math.sqrt(dh)
        # RoPE + KV append in the q/k/v projection's epilogue (one head per 128 x dh output tile) once that grid fills the GPU;
        # few-token prefills (latency path) keep narrower GEMM tiles and the separate RoPE / append launch
        fused = self.fused_qkv_rope
        if fused is None:
            fused = dh in (64, 256) and ((T + 127) // 128) * (Hq + 2 * Hkv) >= 96
        if fused:
            # q, k, v are column slices of `qkv` (token pitch W): the attention's tensor maps take the strides
            q, k, v = qkv, qkv[:, Hq * dh:], qkv[:, (Hq + Hkv) * dh:]
            q_ts = kv_ts = W
        else:
            q = torch.empty(T, Hq * dh, device=dev, dtype=torch.bfloat16)
            k = torch.empty(T, Hkv * dh, device=dev, dtype=torch.bfloat16)
            v = torch.empty(T, Hkv * dh, device=dev, dtype=torch.bfloat16)
            q_ts, kv_ts = Hq * dh, Hkv * dh
        for li, lw in enumerate(pk["layers"]):
            _lib.rmsnorm(h, lw["ln1"], hn)
            kp = kv_cache.k_pages[li].data_ptr() if have_cache else 0
            vp = kv_cache.v_pages[li].data_ptr() if have_cache else 0
            pt = page_table.data_ptr() if have_cache else 0
            sb = slot_base.data_ptr() if have_cache else 0
            mp = page_table.shape[1] if have_cache else 0
            if fused:
                _lib.check(L.pg_gemm_qkv_rope(hn.data_ptr(), D, lw["qkv_w"].data_ptr(), D, qkv.data_ptr(), W, T, D, Hq, Hkv, dh,
                                              pos.data_ptr(), pk["inv_freq"].data_ptr(), kp, vp, pt, sb, S, PAGE, mp, st),
                           "pg_gemm_qkv_rope")
            else:
                _lib.gemm(hn, lw["qkv_w"], qkv, mode=_lib.EPI_BF16, swap=0 if T > 128 else 1)
                _lib.check(L.pg_rope_kv_append(qkv.data_ptr(), 0, pos.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), kp, vp, pt, sb,
                                               B, S, Hq, Hkv, dh, PAGE, mp, pk["inv_freq"].data_ptr(), st), "pg_rope_kv_append")
            # MQA/GQA: the G query heads of a KV head are consecutive rows of one attention problem (no repeat_kv)
            if lens is None:
                _lib.check(L.pg_attention_prefill(
                    q.data_ptr(), k.data_ptr(), v.data_ptr(), att.data_ptr(), B, Hkv, S * G, S, dh, G,
                    S * q_ts, q_ts, dh, G * dh, S * kv_ts, kv_ts, dh,
                    S * Hq * dh, Hq * dh, dh, G * dh, scale, st), "pg_attention_prefill")
            else:
                _lib.check(L.pg_attention_prefill_varlen(
                    q.data_ptr(), k.data_ptr(), v.data_ptr(), att.data_ptr(), lens.data_ptr(), B, Hkv, S * G, S, dh, G,
                    S * q_ts, q_ts, dh, G * dh, S * kv_ts, kv_ts, dh,
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
        caller.  `qkv` is the split-K accumulator of the q/k/v projection: zero on entry of decode_layers and zero again
        on exit (each layer's o_proj launch resets it once the attention kernel has consumed it)."""
        c = self.text_config
        if B > MAX_DECODE_BATCH:
            raise ValueError(f"decode batch {B} > {MAX_DECODE_BATCH}: the weight-streaming decode GEMMs put the batch on the "
                             f"UMMA N axis (<= {MAX_DECODE_BATCH} rows per step); shard the requests over more GPUs / batchers")
        D, F, Hq, Hkv, dh, V = c.hidden_size, c.intermediate_size, c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.vocab_size
        shapes = dict(h=((B, D), torch.float32), hn=((B, D), torch.bfloat16), qkv=((B, (Hq + 2 * Hkv) * dh), torch.float32),
                      att=((B, Hq * dh), torch.bfloat16), mid=((B, F), torch.bfloat16), logits=((B, V), torch.float32),
                      # lm_head epilogue: (max, sum exp2) of every 32-token vocabulary segment, read by the samplers
                      stats=((4 * ((V + 127) // 128), B, 2), torch.float32))

        def make(shp, dt, name):
            return (torch.zeros if name == "qkv" else torch.empty)(*shp, device="cuda", dtype=dt)

        if private:
            return {k: make(shp, dt, k) for k, (shp, dt) in shapes.items()}
        out = {}
        for k, (shp, dt) in shapes.items():
            t = self._ws.get(f"d_{k}_{B}")
            if t is None or t.shape != tuple(shp) or t.dtype != dt:
                t = make(shp, dt, k)
                self._ws[f"d_{k}_{B}"] = t
            out[k] = t
        return out

    # decode-step L2 weight prefetch (decode_layers): bytes of each layer's gate||up weights pulled into L2 during its
    # attention block (0 disables: A/B runs)
    l2_prefetch_bytes = int(float(os.environ.get("PG_L2_PREFETCH_MB", "64")) * 1e6)

    def _prefetch_stream(self):
        s = getattr(self, "_pf_stream", None)
        if s is None or s.device != torch.device("cuda", torch.cuda.current_device()):
            s = self._pf_stream = torch.cuda.Stream()
        return s

    @torch.no_grad()
    def decode_layers(self, bufs, kv_cache: KVCache, B, inv_temperature: float = 1.0):
        """One decode step over all layers (modeling_gemma.py:385-418 at q_len == 1); reads bufs['h'] (fp32 embeddings),
        leaves fp32 logits in bufs['logits'].  Seven launches per layer: RMSNorm, QKV GEMM, fused RoPE + KV append +
        attention, O GEMM, RMSNorm, gate||up GEGLU GEMM, down GEMM (+ one L2 weight prefetch on a forked stream, see below); the three small-output GEMMs split K over CTAs and
        red.add their fp32 partials straight into the residual stream / the zeroed `qkv` accumulator, which the o_proj launch
        resets for the next layer.  (Folding the norms into the GEMMs -- the consumer converting the fp32 residual stream into its
        own bf16 operand tiles -- was built and measured: slower)
        Then the final RMSNorm and the lm_head (modeling_gemma.py:523-525), whose epilogue also leaves the softmax statistics
        of every 32-token vocabulary segment at `inv_temperature` in bufs['stats'] (pg_sample_top_p_stats / pg_argmax_stats).
        Every launch reads its sizes from device counters, so the sequence can be captured in a CUDA graph."""
        c = self.text_config
        pk = self._packed or self.pack()
        L, st = _lib.lib(), _lib.stream()
        D, F, Hq, Hkv, dh, V = c.hidden_size, c.intermediate_size, c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.vocab_size
        h, hn, qkv, att, mid = bufs["h"], bufs["hn"], bufs["qkv"], bufs["att"], bufs["mid"]
        pos, kvl = kv_cache.counters[0], kv_cache.counters[2]
        max_pages = kv_cache.page_table.shape[1]
        scale = 1.0 / math.sqrt(dh)
        eps = c.rms_norm_eps if c.rms_norm_eps is not None else 1e-6
        W = (Hq + 2 * Hkv) * dh
        sms = _lib.num_sms()
        sp_qkv = _pick_split((W + 127) // 128, D // 64, sms)
        sp_o = _pick_split((D + 127) // 128, D // 64, sms)
        sp_down = _pick_split((D + 127) // 128, F // 64, 2 * sms)
        if getattr(self, "deterministic_decode", False):
            # one CTA per output tile: the fp32 red.add has a single writer per element, so the step is bitwise
            # reproducible run to run (the split-K partials otherwise arrive in a different order every launch)
            sp_qkv = sp_o = sp_down = 1
        # L2 weight prefetch on a forked branch (pg_prefetch_l2): while the attention block of a layer (q/k/v, attention, o_proj,
        # norm: ~17 us of kernel-boundary latency with the HBM pins idle) runs, the bulk-copy engines of a few SMs pull the
        # head of the layer's gate||up weights into L2.  Measured per layer (profiles/r02g_decode_l2_prefetch_sweep.txt):
        # 64 x 324 keys 61.6 -> 58.9 us, 1 x 324 51.9 -> 48.5, 32 x 1092 63.9 -> 61.2, 8 x 4164 63.6 -> 63.1.  The stream
        # must not queue ahead of the loads on the critical path: it is forked AFTER the q/k/v launch (64 MB, 32 SMs) when
        # the KV pages of a layer are small, and after the attention launch (48 MB, 64 SMs) when attention itself streams
        # tens of MB; joined once at the end of the step.
        pf_late = B * kv_cache.capacity * Hkv * dh * 4 > 32e6
        pf_bytes = min(self.l2_prefetch_bytes * 3 // 4 if pf_late else self.l2_prefetch_bytes, 2 * F * D * 2) & ~15
        pf_ctas = 64 if pf_late else 32
        cur = torch.cuda.current_stream()
        side = self._prefetch_stream() if pf_bytes >= 16 else None

        def prefetch(lw):
            side.wait_stream(cur)
            _lib.check(L.pg_prefetch_l2(lw["gu_w"].data_ptr(), pf_bytes, pf_ctas, 0, side.cuda_stream), "pg_prefetch_l2")

        for li, lw in enumerate(pk["layers"]):
            _lib.rmsnorm(h, lw["ln1"], hn, eps=eps)
            _lib.gemm(hn, lw["qkv_w"], qkv, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sp_qkv)
            if side is not None and not pf_late:
                prefetch(lw)
            _lib.check(L.pg_attention_decode_fused(
                qkv.data_ptr(), pos.data_ptr(), kvl.data_ptr(), pk["inv_freq"].data_ptr(), kv_cache.k_pages[li].data_ptr(),
                kv_cache.v_pages[li].data_ptr(), kv_cache.page_table.data_ptr(), att.data_ptr(), B, Hq, Hkv, dh, PAGE,
                kv_cache.k_pages.shape[1], max_pages, scale, st), "pg_attention_decode_fused")
            if side is not None and pf_late:
                prefetch(lw)
            _lib.gemm_fused(att, lw["o_w"], h, mode=_lib.EPI_ATOMIC_F32, split_k=sp_o, zero_buf=qkv)
            _lib.rmsnorm(h, lw["ln2"], hn, eps=eps)
            _lib.gemm(hn, lw["gu_w"], mid, mode=_lib.EPI_GEGLU, swap=1)
            _lib.gemm(mid, lw["down_w"], h, mode=_lib.EPI_ATOMIC_F32, swap=1, split_k=sp_down)
        if side is not None:
            cur.wait_stream(side)  # join (the last prefetch finished a layer ago)
        _lib.rmsnorm(h, pk["norm_w"], hn, eps=eps)
        _lib.gemm_fused(hn, pk["head_w"], bufs["logits"], mode=_lib.EPI_F32, bias=pk["head_b"], stats=bufs["stats"],
                        inv_temperature=inv_temperature)
        return bufs["logits"]

    @torch.no_grad()
    def decode_step(self, bufs, kv_cache: KVCache, B, tokens_i32, img, img_scale, pad_token, image_token, inv_temperature=1.0):
        """Embeds `tokens_i32` [B] and runs every layer + final norm + lm_head; fp32 logits land in bufs['logits']."""
        c = self.text_config
        pk = self._packed or self.pack()
        D = c.hidden_size
        _lib.check(_lib.lib().pg_embed_tokens(tokens_i32.data_ptr(), pk["embed"].data_ptr(), _lib.ptr(img), bufs["h"].data_ptr(), B, D,
                                              0 if img is None else img.shape[1], D ** 0.5, img_scale, pad_token, image_token,
                                              _lib.stream()), "pg_embed_tokens")
        return self.decode_layers(bufs, kv_cache, B, inv_temperature)

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
            logits = self.decode_layers(bufs, kv_cache, B).clone().view(B, 1, -1)
            kv_cache._set_len(n + 1, self.text_config.num_hidden_layers)
        else:
            logits = self.prefill(h, pos, B, S, kv_cache, last_only=False)
        out = {"logits": logits}
        if kv_cache is not None:
            out["kv_cache"] = kv_cache
        return out
