"""PaliGemmaForConditionalGeneration with the reference's API (modeling_paligemma.py:14-307) on sm_100a kernels.

forward(input_ids, pixel_values, attention_mask, kv_cache) -> {"logits": fp32 [B,S,V], "kv_cache": KVCache}

Prefill (kv_cache is None or empty): SigLIP tower -> projector GEMM -> fused embedding-gather/image-merge/position-id
kernel -> Gemma prefill.  Decode (q_len == 1): the vision tower is NOT re-run (the reference re-runs it and throws the
result away, modeling_paligemma.py:281-282; skipping it is output-identical) -> token embedding -> Gemma decode step.
`generate()` is the batched, device-side version of the inference.py loop (CUDA-graph replay of one decode step).
"""
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .modeling_gemma import MAX_DECODE_BATCH, GemmaConfig, GemmaForCausalLM, KVCache
from .modeling_siglip import SiglipVisionConfig, SiglipVisionModel, _ParamsOnly, _bf16


class PaliGemmaConfig:
    """Same keyword arguments, defaults and derived fields as modeling_paligemma.py:14-44."""

    def __init__(self, vision_config=None, text_config=None, projection_dim=2048, ignore_index=-100,
                 image_token_index=256000, pad_token_id=None, vocab_size=257152, hidden_size=2048, **kwargs):
        self.projection_dim = projection_dim
        self.ignore_index = ignore_index
        self.image_token_index = image_token_index
        self.pad_token_id = pad_token_id
        self.hidden_size = hidden_size
        self.vision_config = SiglipVisionConfig(**vision_config)
        self.text_config = GemmaConfig(**text_config, pad_token_id=self.pad_token_id)
        self.vocab_size = self.text_config.vocab_size
        self.text_config.num_image_tokens = (self.vision_config.image_size // self.vision_config.patch_size) ** 2
        self.vision_config.projection_dim = projection_dim
        # the reference's loop reads this (inference.py:134); its own config leaves it None unless given
        if self.vision_config.num_image_tokens is None:
            self.vision_config.num_image_tokens = self.text_config.num_image_tokens


class PaliGemmaMultiModalProjector(_ParamsOnly):
    def __init__(self, config: PaliGemmaConfig, **fk):
        super().__init__()
        self.linear = nn.Linear(config.vision_config.hidden_size, config.projection_dim, bias=False, **fk)


class PaliGemmaForConditionalGeneration(nn.Module):
    def __init__(self, config: PaliGemmaConfig, device=None, dtype=None):
        super().__init__()
        fk = {k: v for k, v in dict(device=device, dtype=dtype).items() if v is not None}
        self.config = config
        self.vision_config = config.vision_config
        self.text_config = config.text_config
        if config.projection_dim != config.text_config.hidden_size:
            raise ValueError("projection_dim must equal the text hidden_size (image features are merged into text embeddings)")
        self.vision_tower = SiglipVisionModel(self.vision_config, device=device, dtype=dtype)
        self.language_model = GemmaForCausalLM(self.text_config, device=device, dtype=dtype)
        self.multi_modal_projector = PaliGemmaMultiModalProjector(config, **fk)
        self.pad_token_id = config.pad_token_id if config.pad_token_id is not None else -1
        self.dummy_image_token_id = config.image_token_index
        self._proj_w = None
        self._proj_b = None  # fp32 projector bias of real checkpoints (utils.load_hf_model); the reference has none
        self._graphs = {}

    def tie_weights(self):
        # captured graphs hold raw pointers into the packed weights, which are rebuilt after this
        self._graphs = {}
        self._weights_version = getattr(self, "_weights_version", 0) + 1
        return self.language_model.tie_weights()

    def set_projector_bias(self, bias):
        """multi_modal_projector.linear.bias of a Hugging Face checkpoint (absent from the reference's module tree,
        modeling_paligemma.py:57): added in the projector GEMM epilogue."""
        self._proj_b = None if bias is None else bias.detach().to(device="cuda", dtype=torch.float32).contiguous()
        self._graphs = {}

    def _apply(self, fn, *a, **k):
        self._proj_w = None
        self._graphs = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._proj_w = None
        self._graphs = {}
        self.vision_tower._packed = None
        self.language_model._packed = None
        return super().load_state_dict(*a, **k)

    def pack(self):
        """Packs every weight for the kernels (bf16, fused layouts).  Called lazily by forward/generate."""
        self.vision_tower.pack()
        self.language_model.pack()
        self._proj_w = _bf16(self.multi_modal_projector.linear.weight)
        self._graphs = {}
        self._weights_version = getattr(self, "_weights_version", 0) + 1
        return self

    # -- pieces ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def image_features(self, pixel_values):
        """vision tower + bias-free projector (modeling_paligemma.py:60-65,281-282) -> fp32 [B, N, D] (unscaled)."""
        if self._proj_w is None:
            self._proj_w = _bf16(self.multi_modal_projector.linear.weight)
        B = pixel_values.shape[0]
        feats = self.vision_tower.forward_features(pixel_values, out_bf16=True)
        out = torch.empty(feats.shape[0], self.config.projection_dim, device=feats.device, dtype=torch.float32)
        _lib.gemm(feats, self._proj_w, out, mode=_lib.EPI_F32, bias=self._proj_b, swap=0 if feats.shape[0] > 128 else 1)
        return out.view(B, -1, self.config.projection_dim)

    @torch.no_grad()
    def _merge(self, input_ids, attention_mask, img, validate=True):
        """_merge_input_ids_with_image_features (modeling_paligemma.py:201-251) + the sqrt(D) normaliser of
        modeling_gemma.py:510-511 in one kernel pair.  Returns h fp32 [B*S, D], pos int32 [B*S].  validate=False skips the
        two host-synchronising checks (CUDA-graph capture) and also returns the device error flag for the caller to read."""
        pk = self.language_model._packed or self.language_model.pack()
        B, S = input_ids.shape
        D = self.text_config.hidden_size
        N = img.shape[1]
        dev = img.device
        ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
        mask = attention_mask.to(device=dev).to(torch.int64).contiguous()
        V = self.text_config.vocab_size
        if validate and bool(((ids < 0) | (ids >= V)).any()):
            raise IndexError("input_ids out of range for the embedding table")
        h = torch.empty(B * S, D, device=dev, dtype=torch.float32)
        pos = torch.empty(B * S, device=dev, dtype=torch.int32)
        src = torch.empty(B * S, device=dev, dtype=torch.int32)
        err = torch.zeros(1, device=dev, dtype=torch.int32)
        img_scale = (self.config.projection_dim ** -0.5) * (D ** 0.5)
        _lib.check(_lib.lib().pg_merge_embeddings(
            ids.data_ptr(), mask.data_ptr(), pk["embed"].data_ptr(), img.data_ptr(), h.data_ptr(), pos.data_ptr(),
            src.data_ptr(), err.data_ptr(), B, S, D, N, self.dummy_image_token_id, self.pad_token_id, D ** 0.5, img_scale,
            _lib.stream()), "pg_merge_embeddings")
        if not validate:
            return h, pos, err
        if int(err.item()) != 0:
            raise ValueError(f"every row of input_ids must hold exactly {N} image tokens (id {self.dummy_image_token_id})")
        return h, pos

    @torch.no_grad()
    def _embed_prompt(self, input_ids, attention_mask, pixel_values, validate=True, first_image_row=None):
        """Prefill inputs of the decoder in one pass: vision tower -> projector -> merged, scaled embeddings.
        _merge_input_ids_with_image_features (modeling_paligemma.py:201-251) runs as two halves AROUND the projector GEMM: a scan
        of the ids (position ids, per-token source, the merged row of every image feature) first, then the projector
        (modeling_paligemma.py:57-65) writes its rows, times hidden_size**-0.5 * sqrt(hidden) (:288 and modeling_gemma.py:510-511),
        straight to their `<image>` positions through its epilogue's row map, and a gather fills the text / pad rows.  No fp32
        [B, N, D] image-feature tensor exists.  Returns h fp32 [B*S, D], pos int32 [B*S] (+ the device error flag when
        validate=False: CUDA-graph capture).  first_image_row [B, 1, D] receives every request's first projected, scaled feature row
        (what a decode step needs when the sampled token is `<image>`, modeling_paligemma.py:116-121)."""
        L, lm = _lib.lib(), self.language_model
        pk = lm._packed or lm.pack()
        if self._proj_w is None:
            self._proj_w = _bf16(self.multi_modal_projector.linear.weight)
        B, S = input_ids.shape
        D, V = self.text_config.hidden_size, self.text_config.vocab_size
        dev = torch.device("cuda")
        ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
        mask = attention_mask.to(device=dev).to(torch.int64).contiguous()
        if validate and bool(((ids < 0) | (ids >= V)).any()):
            raise IndexError("input_ids out of range for the embedding table")
        feats = self.vision_tower.forward_features(pixel_values, out_bf16=True)  # [B*N, Dv] bf16
        N = feats.shape[0] // B
        h = torch.empty(B * S + 1, D, device=dev, dtype=torch.float32)  # (+ one sink row for features without a slot)
        pos = torch.empty(B * S, device=dev, dtype=torch.int32)
        src = torch.empty(B * S, device=dev, dtype=torch.int32)
        dst = torch.full((B * N,), B * S, device=dev, dtype=torch.int32)
        err = torch.zeros(1, device=dev, dtype=torch.int32)
        _lib.check(L.pg_merge_scan(ids.data_ptr(), mask.data_ptr(), pos.data_ptr(), src.data_ptr(), dst.data_ptr(), err.data_ptr(), B, S, N,
                                   self.dummy_image_token_id, self.pad_token_id, _lib.stream()), "pg_merge_scan")
        img_scale = (self.config.projection_dim ** -0.5) * (D ** 0.5)
        _lib.gemm(feats, self._proj_w, h, mode=_lib.EPI_F32, bias=self._proj_b, scale=img_scale, swap=0,
                  out_row_map=dst)
        h = h[: B * S]
        _lib.check(L.pg_merge_text(ids.data_ptr(), src.data_ptr(), pk["embed"].data_ptr(), h.data_ptr(), B, S, D, N, D ** 0.5,
                                   _lib.stream()), "pg_merge_text")
        if first_image_row is not None:
            first_image_row.view(B, D).copy_(h.index_select(0, dst.view(B, N)[:, 0].clamp(max=B * S - 1).long()))
        if not validate:
            return h, pos, err
        if int(err.item()) != 0:
            raise ValueError(f"every row of input_ids must hold exactly {N} image tokens (id {self.dummy_image_token_id})")
        return h, pos

    @torch.no_grad()
    def _decode_step(self, tokens_i32, kv_cache, bufs, B, inv_temperature=1.0):
        D = self.text_config.hidden_size
        # image_feats rows written by _embed_prompt already carry the scale
        scale = 1.0 if getattr(kv_cache, "image_feats_scaled", False) else (self.config.projection_dim ** -0.5) * (D ** 0.5)
        return self.language_model.decode_step(bufs, kv_cache, B, tokens_i32, kv_cache.image_feats, scale, self.pad_token_id,
                                               self.dummy_image_token_id, inv_temperature)

    def _sample(self, logits, out_i32, B, do_sample, inv_t, top_p, seed, step, stats=None, seed_dev=None):
        """Next token of every row (inference.py:63-68): greedy argmax or temperature + top-p.  `stats` = the lm_head
        epilogue's segment statistics of these logits at the same inverse temperature (decode steps), else the samplers
        make their own passes over the row (prefill logits)."""
        L, V = _lib.lib(), self.text_config.vocab_size
        if do_sample and stats is not None:
            _lib.check(L.pg_sample_top_p_stats(logits.data_ptr(), logits.stride(0), stats.data_ptr(), stats.shape[1], out_i32.data_ptr(),
                                               B, V, inv_t, float(top_p), int(seed), _lib.ptr(seed_dev), _lib.ptr(step), _lib.stream()),
                       "pg_sample_top_p_stats")
        elif do_sample:
            _lib.check(L.pg_sample_top_p(logits.data_ptr(), logits.stride(0), out_i32.data_ptr(), 0, B, V, inv_t, float(top_p), int(seed),
                                         _lib.ptr(step), _lib.stream()), "pg_sample_top_p")
        elif stats is not None:
            _lib.check(L.pg_argmax_stats(logits.data_ptr(), logits.stride(0), stats.data_ptr(), stats.shape[1], out_i32.data_ptr(), B, V,
                                         _lib.stream()), "pg_argmax_stats")
        else:
            _lib.check(L.pg_argmax(logits.data_ptr(), logits.stride(0), out_i32.data_ptr(), B, V, _lib.stream()), "pg_argmax")

    # -- reference forward ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, input_ids: torch.LongTensor = None, pixel_values: torch.FloatTensor = None,
                attention_mask: Optional[torch.Tensor] = None, kv_cache: Optional[KVCache] = None, last_only: bool = False):
        _lib.require_device()
        B, S = input_ids.shape
        lm = self.language_model
        c = self.text_config
        decode = kv_cache is not None and kv_cache.num_items() > 0
        if not decode:
            first = torch.empty(B, 1, c.hidden_size, device="cuda", dtype=torch.float32) if kv_cache is not None else None
            h, pos = self._embed_prompt(input_ids, attention_mask, pixel_values, first_image_row=first)
            if kv_cache is not None:
                kv_cache.image_feats, kv_cache.image_feats_scaled = first, True
            logits = lm.prefill(h, pos, B, S, kv_cache, last_only=last_only)
        else:
            assert S == 1, "Generation Phase more than one token CAN'T be input"
            n = kv_cache.num_items()
            kv_cache.ensure_capacity(n + 1)
            # position of the new token = number of ones in the grown mask (modeling_paligemma.py:189), per row
            pos = attention_mask.to("cuda").sum(-1).to(torch.int32)
            kv_cache.counters[0].copy_(pos)
            kv_cache.counters[1].fill_(n)
            kv_cache.counters[2].fill_(n + 1)
            bufs = lm.decode_buffers(B)
            tok = input_ids.to(device="cuda", dtype=torch.int32).reshape(B).contiguous()
            logits = self._decode_step(tok, kv_cache, bufs, B).clone().view(B, 1, -1)
            kv_cache._set_len(n + 1, c.num_hidden_layers)
        out = {"logits": logits}
        if kv_cache is not None:
            out["kv_cache"] = kv_cache
        return out

    # -- batched device-side generation (inference.py:45-79 for B rows) --------------------------------------------------
    def _gen_state(self, key, B, S, T, V):
        """Static device buffers (+ the captured decode-step graph) for one generation geometry; reused across calls so
        that graph capture is paid once per (batch, lengths, sampling) configuration."""
        stt = self._graphs.get(key)
        if stt is None:
            dev = torch.device("cuda")
            c = self.text_config
            kv = KVCache(reserve_tokens=T + 1)
            kv.allocate(B, c.num_hidden_layers, c.num_key_value_heads, c.head_dim, S + T + 1)
            stt = dict(kv=kv, nxt=torch.empty(B, device=dev, dtype=torch.int32), cur=torch.empty(B, device=dev, dtype=torch.int32),
                       hist=torch.zeros(T, B, device=dev, dtype=torch.int32), step=torch.zeros(1, device=dev, dtype=torch.int32),
                       seed=torch.zeros(1, device=dev, dtype=torch.int64),
                       img=torch.empty(B, 1, c.hidden_size, device=dev, dtype=torch.float32), graph=None, graph_k=None)
            if len(self._graphs) >= 4:  # bound the number of cached geometries (each owns a KV cache)
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = stt
        return stt

    def _prefill_graphed(self, stt, kv, input_ids, pixel_values, attention_mask, B, S, V):
        """Prefill of generate() through a CUDA graph (captured at the second call of a geometry; the first runs eagerly so
        that lazy kernel attributes and the KV allocation exist).  Input validation keeps the eager path's errors: the id
        range is checked before the launch, the image-token count flag right after the replay."""
        lm, dev = self.language_model, torch.device("cuda")
        ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
        if bool(((ids < 0) | (ids >= V)).any()):
            raise IndexError("input_ids out of range for the embedding table")
        N = self.text_config.num_image_tokens

        def run(ids_t, mask_t, px_t):
            h, pos, err = self._embed_prompt(ids_t, mask_t, px_t, validate=False, first_image_row=stt["img"])
            kv.image_feats, kv.image_feats_scaled = stt["img"], True
            return lm.prefill(h, pos, B, S, kv, last_only=True).view(B, V), err

        pf = stt.get("prefill")
        if pf is None:  # first call: eager (warm-up), remember the static buffers
            logits, err = run(ids, attention_mask.to(dev).to(torch.int64).contiguous(), pixel_values)
            stt["prefill"] = dict(graph=None, ids=torch.empty_like(ids), mask=torch.empty(B, S, device=dev, dtype=torch.int64),
                                  px=torch.empty_like(pixel_values, dtype=torch.float32).contiguous())
        else:
            pf["ids"].copy_(ids)
            pf["mask"].copy_(attention_mask.to(dev).to(torch.int64))
            pf["px"].copy_(pixel_values)
            if pf["graph"] is None:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        pf["logits"], pf["err"] = run(pf["ids"], pf["mask"], pf["px"])
                torch.cuda.current_stream().wait_stream(side)
                pf["graph"] = g
            pf["graph"].replay()
            logits, err = pf["logits"], pf["err"]
        if int(err.item()) != 0:
            raise ValueError(f"every row of input_ids must hold exactly {N} image tokens (id {self.dummy_image_token_id})")
        return logits

    @torch.no_grad()
    def generate(self, input_ids, pixel_values, attention_mask, max_tokens_to_generate: int, do_sample: bool = False,
                 temperature: float = 0.8, top_p: float = 0.9, eos_token_id: Optional[int] = None, seed: int = 0,
                 use_cuda_graph: bool = True, return_logits: bool = False, forced_tokens=None, timings: dict = None,
                 prompt_lens=None):
        """Returns int64 tokens [B, T] (T = max_tokens_to_generate, or shorter if every row has emitted EOS; rows that
        finished early keep generating -- trim at the first EOS as the reference loop would).  The decode step
        (embedding, all layers, lm_head, sampler, counter advance) is captured once in a CUDA graph and replayed; the host
        only checks EOS every 16 steps.  `forced_tokens` [B, T] teacher-forces the step inputs (parity tests).

        `prompt_lens` [B] serves a RAGGED batch: row b's prompt is its first prompt_lens[b] tokens, the rest of the row is
        right padding that is masked out (each row then gets exactly the result of its own B = 1 run).  Without it a
        padded row behaves as in the reference: pads embedded as zeros at position 1 and attended to
        (modeling_paligemma.py:125-127,154-156,195)."""
        _lib.require_device()
        L = _lib.lib()
        lm, c = self.language_model, self.text_config
        B, S = input_ids.shape
        if B > MAX_DECODE_BATCH:
            raise ValueError(f"generate() serves at most {MAX_DECODE_BATCH} rows per call (decode batch of the weight-streaming "
                             "GEMMs); split the batch or shard it over GPUs (sharding.py)")
        T = int(max_tokens_to_generate)
        V = c.vocab_size
        dev = torch.device("cuda")
        key = (B, S, T, bool(do_sample), float(temperature), float(top_p))  # (the seed lives in device memory)
        stt = self._gen_state(key, B, S, T, V)
        kv, nxt, cur, hist, step = stt["kv"], stt["nxt"], stt["cur"], stt["hist"], stt["step"]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timings is not None else None
        if ev:
            ev[0].record()
        lens = None
        if prompt_lens is not None:
            lens = torch.as_tensor(prompt_lens).to(device=dev, dtype=torch.int32).reshape(B).contiguous()
            if bool(((lens < 1) | (lens > S)).any()):
                raise ValueError("prompt_lens must lie in [1, S]")
            attention_mask = (torch.arange(S, device=dev)[None, :] < lens[:, None]).to(torch.int64)
        # Latency path (few tokens: the ~350 prefill launches are bound by the host's enqueue rate, 4.3 ms of host time for
        # 5.1 ms of prefill at one request): from the second call of a geometry on, the whole prefill -- vision tower,
        # projector, merge, decoder, last-position head -- replays as ONE CUDA graph over static input buffers.
        graph_prefill = use_cuda_graph and lens is None and B * S <= 1100 and pixel_values.is_cuda and attention_mask is not None
        if graph_prefill:
            logits = self._prefill_graphed(stt, kv, input_ids, pixel_values, attention_mask, B, S, V)
        else:
            h, pos = self._embed_prompt(input_ids, attention_mask, pixel_values, first_image_row=stt["img"])  # static: the decode graph reads it
            kv.image_feats, kv.image_feats_scaled = stt["img"], True
            logits = lm.prefill(h, pos, B, S, kv, last_only=True, lens=lens).view(B, V)
        if ev:
            ev[1].record()

        bufs = lm.decode_buffers(B)
        logit_log = torch.empty(T, B, V, device=dev, dtype=torch.float32) if return_logits else None
        forced = None if forced_tokens is None else forced_tokens.to(device=dev, dtype=torch.int32).t().contiguous()
        # decode counters: position id of the next token, its cache slot, kv length including it
        kv.counters[0].copy_(attention_mask.to(dev).sum(-1).to(torch.int32) + 1)
        if lens is None:
            kv.counters[1].fill_(S)
            kv.counters[2].fill_(S + 1)
        else:  # ragged: row b continues right after its own prompt (its padding slots are overwritten as it grows)
            kv.counters[1].copy_(lens)
            kv.counters[2].copy_(lens + 1)
        step.zero_()
        inv_t = 1.0 / float(temperature) if do_sample else 1.0
        stt["seed"].fill_(int(seed))

        def sample(lg, stats=None):
            self._sample(lg, nxt, B, do_sample, inv_t, top_p, seed, step, stats=stats, seed_dev=stt["seed"])

        def advance(n_counters):
            _lib.check(L.pg_advance_decode(nxt.data_ptr(), hist.data_ptr(), cur.data_ptr(), kv.counters.data_ptr(), n_counters,
                                           step.data_ptr(), B, _lib.stream()), "pg_advance_decode")

        def decode_step(t=None):
            lg = self._decode_step(cur, kv, bufs, B, inv_t)
            if t is not None and return_logits:
                logit_log[t].copy_(lg)
            sample(lg, bufs["stats"])
            advance(3)
            if t is not None and forced is not None:
                cur.copy_(forced[t])

        # token 0 comes from the prefill logits
        if return_logits:
            logit_log[0].copy_(logits)
        sample(logits)
        advance(0)
        if forced is not None:
            cur.copy_(forced[0])

        graphed = use_cuda_graph and not return_logits and forced is None
        done_at = T
        GK = 8  # decode steps per replay of the multi-step graph (amortises the graph-launch latency)

        def capture(n_steps):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for _ in range(n_steps):
                        decode_step()
            torch.cuda.current_stream().wait_stream(side)
            return g

        t = 1  # tokens generated so far
        while t < T:
            t_prev = t
            if graphed and (t >= 2 or stt["graph"] is not None):
                if stt["graph"] is None:  # step 1 ran eagerly (lazy kernel attributes / driver entry points are warm)
                    stt["graph"] = capture(1)  # (the capture itself does not execute)
                    stt["graph_k"] = capture(GK) if T - 2 >= GK else None
                if T - t >= GK and stt.get("graph_k") is not None:
                    stt["graph_k"].replay()
                    t += GK
                else:
                    stt["graph"].replay()
                    t += 1
            else:
                decode_step(t)
                t += 1
            if eos_token_id is not None and (t // 16 > t_prev // 16 or t == T):
                if bool((hist[:t] == eos_token_id).any(0).all()):
                    done_at = t
                    break
        kv._set_len(S + done_at - 1, c.num_hidden_layers)
        if ev:
            ev[2].record()
            torch.cuda.synchronize()
            timings["prefill_ms"] = ev[0].elapsed_time(ev[1])
            timings["decode_ms"] = ev[1].elapsed_time(ev[2])
            timings["decode_steps"] = done_at - 1
        toks = hist[:done_at].t().contiguous().long()
        if return_logits:
            return toks, logit_log[:done_at].transpose(0, 1).contiguous()
        return toks
