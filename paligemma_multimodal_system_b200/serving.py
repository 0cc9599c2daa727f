"""Variable-length prompts and continuous batching on top of the paged KVCache.

The reference loop (inference.py:45-79) serves ONE request (`assert next_token.size() == (1, 1)`, :69): prefill, then one
token per step until EOS has been appended (:71-74) or `max_tokens_to_generate` is reached.  Here B such loops share one
decode batch.

  slot      one row of the decode batch: three device counters (position id, KV write slot, KV length), a KV budget, the
            current token, one row of the LIVE page table.  The decode step reads all of that from device memory, so ONE
            captured CUDA graph of `steps_per_replay` steps serves every mix of requests.
  page set  the pages that hold one request's keys/values (one row of the HOME page table).  There are `stage` more sets
            than slots: queued requests are prefilled AHEAD into free sets, in groups large enough for the tensor-core
            prefill to be efficient, and wait there ("staged") with their first token already sampled.  When a slot
            retires, a staged request takes it over by copying its set's page-table row into the slot's live row -- no
            key/value moves -- so slots are refilled at every replay boundary without paying a tiny prefill.

Between graph replays the host reads the ring of sampled tokens, trims every request at its first EOS / at its token budget,
retires it, arms staged requests into the free slots and, when enough sets are free, runs ONE ragged prefill for the next
group (prompts right-padded to the longest of the group; row b attends to its own lens[b] keys only --
pg_attention_prefill_varlen -- and its first token comes from its last *real* position).  Every request therefore sees
exactly the arithmetic of its own B = 1 run (unpadded prompt, positions 1..S, its own KV length), which is what the parity
tests check against vectors of the unmodified reference and against `generate()` at B = 1.  The reference's own treatment
of right padding (pads zeroed, position 1, never masked: modeling_paligemma.py:125-127,154-156,195) stays available through
`forward()` with an attention_mask.

`SlotScheduler` is the host-side bookkeeping alone (no tensors; CPU-testable); `ContinuousBatcher` is the device engine.
"""
import collections
import dataclasses
import time
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from .modeling_gemma import KVCache


@dataclasses.dataclass
class Request:
    rid: int
    input_ids: torch.Tensor      # int64 [S]: N image tokens followed by the text prompt (processing_paligemma.py:77-89)
    pixel_values: torch.Tensor   # [3, H, W]
    max_new_tokens: int
    tokens: List[int] = dataclasses.field(default_factory=list)
    page_set: int = -1           # row of the home page table that holds this request's keys/values


class SlotScheduler:
    """FIFO queue -> page sets (prefill, possibly ahead of need) -> slots (decode); token accounting with the reference
    loop's stop rule (EOS is appended, then the request stops, inference.py:71-74; otherwise after max_new_tokens)."""

    def __init__(self, num_slots: int, eos_token_id: Optional[int] = None, min_admit: int = 1, stage: int = 0):
        if num_slots <= 0 or stage < 0:
            raise ValueError("num_slots must be positive and stage non-negative")
        self.num_slots = num_slots
        self.eos_token_id = eos_token_id
        self.min_admit = max(1, int(min_admit))
        self.free_slots: List[int] = list(range(num_slots))
        self.free_sets: List[int] = list(range(num_slots + int(stage)))
        self.queue = collections.deque()
        self.staged = collections.deque()
        self.active: Dict[int, Request] = {}
        self.finished: Dict[int, List[int]] = {}

    def submit(self, req: Request):
        self.queue.append(req)

    def idle(self) -> bool:
        return not self.queue and not self.staged and not self.active

    def _stops(self, req: Request, t: int) -> bool:
        return len(req.tokens) >= req.max_new_tokens or (self.eos_token_id is not None and t == self.eos_token_id)

    def plan_prefill(self) -> List[Request]:
        """Requests to prefill now (each gets a free page set, lowest first).  A prefill of very few rows costs as much as
        several decode steps of the whole batch, so while anything is decoding or staged the scheduler waits until
        `min_admit` rows (or the whole remaining queue) can go together."""
        n = min(len(self.free_sets), len(self.queue))
        if n == 0:
            return []
        if (self.active or self.staged) and n < min(self.min_admit, len(self.queue)):
            return []
        self.free_sets.sort()
        reqs = []
        for _ in range(n):
            req = self.queue.popleft()
            req.page_set = self.free_sets.pop(0)
            reqs.append(req)
        return reqs

    def prefilled(self, req: Request, first_token: int):
        """The prefill sampled the request's first token: it is staged, unless that token already ends it."""
        req.tokens.append(int(first_token))
        if self._stops(req, int(first_token)):
            self.finished[req.rid] = req.tokens
            self.free_sets.append(req.page_set)
        else:
            self.staged.append(req)

    def plan_arming(self) -> List[Tuple[int, Request]]:
        """Staged requests (oldest first) take over the free slots (lowest first)."""
        self.free_slots.sort()
        pairs = []
        while self.free_slots and self.staged:
            slot, req = self.free_slots.pop(0), self.staged.popleft()
            self.active[slot] = req
            pairs.append((slot, req))
        return pairs

    def consume(self, slot: int, toks: List[int]) -> bool:
        """Appends the tokens a slot produced (in order); returns True when the request finished (slot and page set freed).
        Tokens after the stop point (the slot keeps stepping until the replay ends) are dropped."""
        req = self.active[slot]
        for t in toks:
            req.tokens.append(int(t))
            if self._stops(req, int(t)):
                self.finished[req.rid] = req.tokens
                del self.active[slot]
                self.free_slots.append(slot)
                self.free_sets.append(req.page_set)
                return True
        return False


class ContinuousBatcher:
    def __init__(self, model, num_slots: int, max_prompt_len: int, max_new_tokens: int, do_sample: bool = False,
                 temperature: float = 0.8, top_p: float = 0.9, eos_token_id: Optional[int] = None, seed: int = 0,
                 steps_per_replay: int = 8, min_admit: int = 1, stage: int = 0, use_cuda_graph: bool = True,
                 keep_admit_logits: bool = False):
        """max_prompt_len counts the image tokens too (S = num_image_tokens + text tokens).  `stage` = page sets beyond the
        slots (requests prefilled ahead); `min_admit` = rows a prefill waits for while the batch is busy."""
        _lib.require_device()
        self.model = model
        self.lm, self.c = model.language_model, model.text_config
        c = self.c
        self.B = int(num_slots)
        from .modeling_gemma import MAX_DECODE_BATCH
        if self.B > MAX_DECODE_BATCH:
            raise ValueError(f"at most {MAX_DECODE_BATCH} slots per batcher (one decode step = one <= {MAX_DECODE_BATCH}-row "
                             "weight-streaming GEMM batch); run several batchers / GPUs for more concurrent requests")
        self.max_prompt_len, self.max_new_tokens = int(max_prompt_len), int(max_new_tokens)
        self.do_sample, self.inv_t, self.top_p, self.seed = bool(do_sample), 1.0 / float(temperature), float(top_p), int(seed)
        self.GK = max(1, int(steps_per_replay))
        self.use_cuda_graph = use_cuda_graph
        self.sched = SlotScheduler(self.B, eos_token_id, min_admit, stage)
        dev = torch.device("cuda")
        n_sets = self.B + int(stage)
        self.scratch_set = n_sets  # idle slots write their (meaningless) keys/values here, never into a request's pages
        self.kv = KVCache()
        self.kv.allocate(n_sets + 1, c.num_hidden_layers, c.num_key_value_heads, c.head_dim, self.max_prompt_len + self.max_new_tokens)
        self.home_table = self.kv.page_table                                      # [n_sets + 1, max_pages]
        self.kv.page_table = self.home_table[: self.B].clone()                    # live rows, one per slot (what decode reads)
        self.kv.counters = torch.zeros(3, self.B, device=dev, dtype=torch.int32)  # per slot
        self.kv._set_len(1, c.num_hidden_layers)  # "decode phase" for anything that asks num_items()
        self.cur = torch.zeros(self.B, device=dev, dtype=torch.int32)
        self.nxt = torch.zeros(self.B, device=dev, dtype=torch.int32)
        self.ring = torch.zeros(self.GK, self.B, device=dev, dtype=torch.int32)
        self.step = torch.zeros(1, device=dev, dtype=torch.int32)
        self.limit = torch.ones(self.B, device=dev, dtype=torch.int32)
        # a decode step only ever consumes the FIRST projected feature row of the request's image (an `<image>` token sampled
        # at q_len = 1, modeling_paligemma.py:116-121): one row per slot / per page set is all that has to be kept
        self.img = torch.zeros(self.B, 1, c.hidden_size, device=dev, dtype=torch.float32)
        self.img_sets = torch.zeros(n_sets, c.hidden_size, device=dev, dtype=torch.float32)
        self.kv.image_feats, self.kv.image_feats_scaled = self.img, True  # rows come from _embed_prompt, already scaled
        self._set_idle(list(range(self.B)))
        self.bufs = self.lm.decode_buffers(self.B, private=True)  # the captured graph owns these addresses
        self.graph = None
        self._graph_weights_version = 0
        self.admit_logits = {} if keep_admit_logits else None  # request id -> fp32 logits of its first token (tests)
        self._next_rid = 0
        self.reset_stats()

    def reset_stats(self):
        self.stats = dict(prefill_groups=0, prefill_rows=0, decode_replays=0, decode_steps=0, tokens=0, wall_s=0.0,
                          t_prefill_s=0.0, t_prefill_prep_s=0.0, t_arm_s=0.0, t_decode_s=0.0, t_retire_s=0.0)  # host wall time per phase

    # -- host API --------------------------------------------------------------------------------------------------
    def submit(self, input_ids: torch.Tensor, pixel_values: torch.Tensor, max_new_tokens: Optional[int] = None) -> int:
        ids = input_ids.reshape(-1).to("cpu", torch.int64)
        m = self.max_new_tokens if max_new_tokens is None else int(max_new_tokens)
        if ids.numel() > self.max_prompt_len or ids.numel() < 1:
            raise ValueError(f"prompt of {ids.numel()} tokens does not fit max_prompt_len = {self.max_prompt_len}")
        if not 1 <= m <= self.max_new_tokens:
            raise ValueError(f"max_new_tokens must lie in [1, {self.max_new_tokens}]")
        # a bad request is refused HERE, never in the middle of a prefill group that carries other requests
        if bool(((ids < 0) | (ids >= self.c.vocab_size)).any()):
            raise IndexError("input_ids out of range for the embedding table")
        n_img = int((ids == self.model.dummy_image_token_id).sum())
        if n_img != self.c.num_image_tokens:
            raise ValueError(f"a request must hold exactly {self.c.num_image_tokens} image tokens (id "
                             f"{self.model.dummy_image_token_id}), got {n_img}")
        vc = self.model.vision_config
        if tuple(pixel_values.shape) != (vc.num_channels, vc.image_size, vc.image_size):
            raise ValueError(f"pixel_values must be [{vc.num_channels}, {vc.image_size}, {vc.image_size}], got {tuple(pixel_values.shape)}")
        rid = self._next_rid
        self._next_rid += 1
        self.sched.submit(Request(rid, ids, pixel_values, m))
        return rid

    @torch.no_grad()
    def run(self) -> Dict[int, torch.Tensor]:
        """Serves every submitted request; returns {request id: int64 tokens} (EOS included when it was emitted)."""
        t0 = time.perf_counter()
        clk, st = time.perf_counter, self.stats
        while not self.sched.idle():
            ta = clk()
            reqs = self.sched.plan_prefill()
            if reqs:
                self._prefill(reqs)  # (ends with the D2H read of the first tokens: device time of the prefill included)
            tb = clk()
            pairs = self.sched.plan_arming()
            if pairs:
                self._arm(pairs)
            tc = clk()
            if self.sched.active:
                self._decode_group()
            st["t_prefill_s"] += tb - ta
            st["t_arm_s"] += tc - tb
            st["t_decode_s"] += clk() - tc
        torch.cuda.synchronize()
        self.stats["wall_s"] += time.perf_counter() - t0
        out = {rid: torch.tensor(t, dtype=torch.int64) for rid, t in self.sched.finished.items()}
        self.sched.finished = {}
        return out

    # -- device side -----------------------------------------------------------------------------------------------
    def _set_idle(self, slots: List[int]):
        if not slots:
            return
        s = torch.tensor(slots, device="cuda", dtype=torch.int64)
        one = torch.ones(len(slots), device="cuda", dtype=torch.int32)
        self.kv.page_table.index_copy_(0, s, self.home_table[self.scratch_set].expand(len(slots), -1).contiguous())
        self.kv.counters[0].index_copy_(0, s, one)      # position id 1
        self.kv.counters[1].index_copy_(0, s, one - 1)  # write slot 0
        self.kv.counters[2].index_copy_(0, s, one)      # one key
        self.limit.index_copy_(0, s, one)               # kv_len == limit: frozen
        self.cur.index_fill_(0, s, 0)

    def _sample(self, logits, out, rows, seed, stats=None):
        self.model._sample(logits, out, rows, self.do_sample, self.inv_t, self.top_p, seed, self.step, stats=stats)

    def _prefill(self, reqs: List[Request]):
        """One ragged prefill: keys/values of request i land in its page set, its first token is sampled."""
        model, c = self.model, self.c
        t_prep = time.perf_counter()
        g = len(reqs)
        lens = [int(r.input_ids.numel()) for r in reqs]
        S = max(lens)
        fill = model.pad_token_id if model.pad_token_id is not None and model.pad_token_id >= 0 else 0
        ids = torch.full((g, S), fill, dtype=torch.int64)
        mask = torch.zeros(g, S, dtype=torch.int64)
        for i, r in enumerate(reqs):
            ids[i, : lens[i]] = r.input_ids
            mask[i, : lens[i]] = 1
        px = torch.stack([r.pixel_values for r in reqs]).to("cuda", non_blocking=True)
        sets_t = torch.tensor([r.page_set for r in reqs], device="cuda", dtype=torch.int64)
        lens_t = torch.tensor(lens, device="cuda", dtype=torch.int32)
        self.stats["t_prefill_prep_s"] += time.perf_counter() - t_prep  # host batching + H2D of the pixels
        first = torch.empty(g, 1, c.hidden_size, device="cuda", dtype=torch.float32)
        h, pos = model._embed_prompt(ids.cuda(), mask.cuda(), px, first_image_row=first)
        self.img_sets.index_copy_(0, sets_t, first[:, 0])
        logits = self.lm.prefill(h, pos, g, S, self.kv, last_only=True, lens=lens_t,
                                 page_rows=self.home_table.index_select(0, sets_t)).view(g, c.vocab_size)
        if self.admit_logits is not None:
            for i, r in enumerate(reqs):
                self.admit_logits[r.rid] = logits[i].clone()
        first = torch.empty(g, device="cuda", dtype=torch.int32)
        self._sample(logits, first, g, self.seed ^ 0x5DEECE66D)  # prefills draw from their own RNG stream
        self.stats["prefill_groups"] += 1
        self.stats["prefill_rows"] += g
        self.stats["tokens"] += g
        for r, t in zip(reqs, first.tolist()):
            self.sched.prefilled(r, t)

    def _arm(self, pairs: List[Tuple[int, Request]]):
        """Staged requests take over slots: page-table row, first image feature row, current token, counters, budget."""
        slots_t = torch.tensor([s for s, _ in pairs], device="cuda", dtype=torch.int64)
        sets_t = torch.tensor([r.page_set for _, r in pairs], device="cuda", dtype=torch.int64)
        lens_t = torch.tensor([int(r.input_ids.numel()) for _, r in pairs], device="cuda", dtype=torch.int32)
        budget = torch.tensor([r.max_new_tokens for _, r in pairs], device="cuda", dtype=torch.int32)
        first = torch.tensor([r.tokens[0] for _, r in pairs], device="cuda", dtype=torch.int32)
        self.kv.page_table.index_copy_(0, slots_t, self.home_table.index_select(0, sets_t))
        self.img[:, 0].index_copy_(0, slots_t, self.img_sets.index_select(0, sets_t))
        self.cur.index_copy_(0, slots_t, first)
        # next token: position id S+1 (1-based positions, modeling_paligemma.py:189), written at cache slot S, kv length S+1
        self.kv.counters[0].index_copy_(0, slots_t, lens_t + 1)
        self.kv.counters[1].index_copy_(0, slots_t, lens_t)
        self.kv.counters[2].index_copy_(0, slots_t, lens_t + 1)
        self.limit.index_copy_(0, slots_t, lens_t + budget - 1)  # the step that produces token max_new-1 is the last to advance

    def _consume(self, slot, toks):
        req = self.sched.active[slot]
        before = len(req.tokens)
        fin = self.sched.consume(slot, toks)
        self.stats["tokens"] += len(req.tokens) - before
        return fin

    def _step(self):
        lg = self.model._decode_step(self.cur, self.kv, self.bufs, self.B, self.inv_t if self.do_sample else 1.0)
        self._sample(lg, self.nxt, self.B, self.seed, stats=self.bufs["stats"])
        _lib.check(_lib.lib().pg_advance_decode_slots(self.nxt.data_ptr(), self.ring.data_ptr(), self.GK, self.cur.data_ptr(),
                                                      self.kv.counters.data_ptr(), self.limit.data_ptr(), self.step.data_ptr(),
                                                      self.B, _lib.stream()), "pg_advance_decode_slots")

    def _capture(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(self.GK):
                    self._step()
        torch.cuda.current_stream().wait_stream(side)
        return g

    def _decode_group(self):
        ver = getattr(self.model, "_weights_version", 0)
        if self.graph is not None and ver != self._graph_weights_version:
            self.graph = None  # the model re-packed its weights (tie_weights / pack): the captured pointers are stale
        if self.use_cuda_graph and self.graph is None and self.stats["decode_replays"] > 0:
            self.graph = self._capture()  # the first group ran eagerly: lazy kernel attributes / entry points are warm
            self._graph_weights_version = ver
        if self.graph is not None:
            self.graph.replay()
        else:
            for _ in range(self.GK):
                self._step()
        ring = self.ring.cpu()  # [GK, B]; the one host sync per group
        t0 = time.perf_counter()
        self.stats["decode_replays"] += 1
        self.stats["decode_steps"] += self.GK
        cols = ring.t().tolist()
        done = [slot for slot in list(self.sched.active) if self._consume(slot, cols[slot])]
        self._set_idle(done)
        self.stats["t_retire_s"] += time.perf_counter() - t0
