"""Variable-length prompts + continuous batching (serving.py) on the GPU, through the C ABI.

Judges: plain PyTorch fp32 for the ragged attention kernel; vectors of the UNMODIFIED reference run one request at a time
(tests/golden/serving_reference.npz) and this package's own B = 1 generate() for the batcher."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paligemma_multimodal_system_b200.random_init import TINY_CONFIG, make_requests, make_state_dict  # noqa: E402
from parity_utils import build_model, stats  # noqa: E402

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, "tests", "golden", "serving_reference.npz"))


@pytest.mark.parametrize("B,H,N,dh,group,lens", [
    (3, 1, 264, 256, 8, [264, 258, 1]), (4, 1, 266, 64, 4, [259, 266, 130, 64]), (2, 16, 300, 72, 1, [300, 129]),
    (2, 1, 1030, 256, 8, [1030, 65])])
def test_prefill_attention_ragged_key_counts(B, H, N, dh, group, lens):
    """pg_attention_prefill_varlen: problem b attends to its first lens[b] keys only; the padding keys hold finite junk."""
    from paligemma_multimodal_system_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(11)
    q = (torch.randn(B, N, H, group, dh, device="cuda", generator=g) * 0.8).bfloat16()
    k = (torch.randn(B, N, H, dh, device="cuda", generator=g) * 0.8).bfloat16()
    v = (torch.randn(B, N, H, dh, device="cuda", generator=g) * 0.8).bfloat16()
    for b, n in enumerate(lens):  # junk that would dominate the softmax if it were not masked
        k[b, n:] *= 20.0
        v[b, n:] = 100.0
    out = torch.full((B, N, H, group, dh), float("nan"), device="cuda", dtype=torch.bfloat16)
    lens_t = torch.tensor(lens, device="cuda", dtype=torch.int32)
    scale = dh ** -0.5
    rc = _lib.lib().pg_attention_prefill_varlen(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lens_t.data_ptr(), B, H, N * group, N, dh, group,
        N * H * group * dh, H * group * dh, dh, group * dh, N * H * dh, H * dh, dh,
        N * H * group * dh, H * group * dh, dh, group * dh, scale, _lib.stream())
    _lib.check(rc, "attn varlen")
    torch.cuda.synchronize()
    for b, n in enumerate(lens):
        qf = q[b, :n].float().permute(1, 2, 0, 3)                 # [H,G,n,dh]
        kf = k[b, :n].float().permute(1, 0, 2).unsqueeze(1)       # [H,1,n,dh]
        vf = v[b, :n].float().permute(1, 0, 2).unsqueeze(1)
        ref = (torch.softmax(qf @ kf.transpose(-1, -2) * scale, -1) @ vf).permute(2, 0, 1, 3)  # [n,H,G,dh]
        got = out[b, :n].float()
        err = (got - ref).abs().max().item()
        assert torch.isfinite(got).all() and err <= 2e-2 * max(ref.abs().max().item(), 1.0), (b, n, err)
    # without key counts the same call must be an argument error, and unsupported head sizes too
    assert _lib.lib().pg_attention_prefill_varlen(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), 0, B, H, N * group, N, dh, group,
        N * H * group * dh, H * group * dh, dh, group * dh, N * H * dh, H * dh, dh,
        N * H * group * dh, H * group * dh, dh, group * dh, scale, _lib.stream()) == -1


def test_advance_decode_slots_freezes_at_budget_and_rings():
    from paligemma_multimodal_system_b200 import _lib
    B, ring = 5, 4
    nxt = torch.zeros(B, device="cuda", dtype=torch.int32)
    cur = torch.full((B,), -1, device="cuda", dtype=torch.int32)
    log = torch.full((ring, B), -1, device="cuda", dtype=torch.int32)
    counters = torch.tensor([[11, 21, 1, 41, 51], [10, 20, 0, 40, 50], [11, 21, 1, 41, 51]], device="cuda", dtype=torch.int32)
    limit = torch.tensor([13, 21, 1, 100, 52], device="cuda", dtype=torch.int32)
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    for s in range(6):
        nxt.copy_(torch.arange(B, device="cuda", dtype=torch.int32) + 100 * s)
        _lib.check(_lib.lib().pg_advance_decode_slots(nxt.data_ptr(), log.data_ptr(), ring, cur.data_ptr(), counters.data_ptr(),
                                                      limit.data_ptr(), step.data_ptr(), B, _lib.stream()), "advance slots")
    torch.cuda.synchronize()
    assert int(step) == 6
    assert counters[2].tolist() == [13, 21, 1, 47, 52]   # slot 0 stops after two advances, 1 and 2 never move, 4 moves once
    assert counters[0].tolist() == [13, 21, 1, 47, 52] and counters[1].tolist() == [12, 20, 0, 46, 51]
    assert cur.tolist() == [500, 501, 502, 503, 504]
    assert log[:, 0].tolist() == [400, 500, 200, 300]    # ring of 4: steps 4, 5 overwrote rows 0, 1
    assert _lib.lib().pg_advance_decode_slots(nxt.data_ptr(), log.data_ptr(), 0, cur.data_ptr(), counters.data_ptr(),
                                              limit.data_ptr(), step.data_ptr(), B, _lib.stream()) == -1


def _serve(model, reqs, budgets, **kw):
    from paligemma_multimodal_system_b200.serving import ContinuousBatcher
    cb = ContinuousBatcher(model, max_prompt_len=max(int(i.numel()) for i, _ in reqs), max_new_tokens=max(budgets), **kw)
    rids = [cb.submit(ids, px, m) for (ids, px), m in zip(reqs, budgets)]
    out = cb.run()
    assert sorted(out) == sorted(rids)
    return [out[r].tolist() for r in rids], cb


@pytest.mark.parametrize("regime", ["R0", "R1"])
@pytest.mark.parametrize("graph,stage,min_admit", [(False, 0, 1), (True, 0, 1), (True, 3, 2), (True, 8, 4)])
def test_continuous_batching_reproduces_the_reference_request_by_request(regime, graph, stage, min_admit):
    """10 ragged requests through 4 slots (so slots are refilled while others are mid-generation), different token budgets:
    every request must produce exactly what the unmodified reference produced for it alone."""
    model = build_model(TINY_CONFIG, make_state_dict(TINY_CONFIG, regime, seed=11))
    reqs = make_requests(TINY_CONFIG, 10, 2, 8, seed=21)
    budgets = [12, 5, 9, 1, 12, 7, 3, 12, 10, 6]
    toks, cb = _serve(model, reqs, budgets, num_slots=4, steps_per_replay=4, use_cuda_graph=graph, stage=stage, min_admit=min_admit)
    for r in range(10):
        assert toks[r] == G[f"{regime}_tokens"][r][: budgets[r]].tolist(), (r, toks[r])
    assert cb.stats["prefill_rows"] == 10 and cb.stats["prefill_groups"] >= (3 if stage == 0 else 1) and cb.stats["tokens"] == sum(budgets)
    assert sorted(cb.sched.free_sets) == list(range(4 + stage)) and sorted(cb.sched.free_slots) == [0, 1, 2, 3]
    # the batcher is reusable (captured graph, slots back to idle): a second wave gives the same answers
    rids = [cb.submit(*reqs[r], 8) for r in (9, 0, 5)]
    out = cb.run()
    for rid, r in zip(rids, (9, 0, 5)):
        assert out[rid].tolist() == G[f"{regime}_tokens"][r][:8].tolist()


@pytest.mark.parametrize("stage", [0, 2])
def test_continuous_batching_stops_at_eos_and_refills(stage):
    """The reference appends EOS and stops (inference.py:71-74).  In R1 request r keeps emitting its own last prompt token,
    so declaring request 1's token the EOS id ends that request after one token while the others run to their budgets."""
    model = build_model(TINY_CONFIG, make_state_dict(TINY_CONFIG, "R1", seed=11))
    reqs = make_requests(TINY_CONFIG, 6, 2, 8, seed=21)
    eos = int(G["R1_tokens"][1][0])
    assert all(eos not in G["R1_tokens"][r].tolist() for r in (0, 2, 3, 4, 5))
    toks, cb = _serve(model, reqs, [10] * 6, num_slots=2, steps_per_replay=3, eos_token_id=eos, stage=stage)
    assert toks[1] == [eos]
    for r in (0, 2, 3, 4, 5):
        assert toks[r] == G["R1_tokens"][r][:10].tolist()


def test_ragged_prefill_logits_and_single_request_runs_R2():
    """Diffuse regime: the first token's logits of every admitted row against the reference's B = 1 prefill logits (stated
    tolerance of the tiny R2 regime: cosine >= 0.999, max-abs <= 3 % of absmax); free-running tokens are compared with
    this package's own B = 1 generate() and reported (greedy in R2 is not stable under bf16 noise)."""
    from paligemma_multimodal_system_b200.serving import ContinuousBatcher
    model = build_model(TINY_CONFIG, make_state_dict(TINY_CONFIG, "R2", seed=11))
    reqs = make_requests(TINY_CONFIG, 10, 2, 8, seed=21)
    cb = ContinuousBatcher(model, num_slots=6, max_prompt_len=264, max_new_tokens=12, keep_admit_logits=True, stage=2, min_admit=2)
    rids = [cb.submit(ids, px) for ids, px in reqs]
    out = cb.run()
    for r in range(10):
        s = stats(cb.admit_logits[rids[r]], torch.as_tensor(G["R2_prefill_logits"][r]))
        assert s["cos"] >= 0.999 and s["rel"] <= 3e-2, (r, s)
    agree = []
    for r in range(10):
        ids, px = reqs[r]
        solo = model.generate(ids[None].cuda(), px[None].cuda(), torch.ones(1, ids.numel(), dtype=torch.int64).cuda(), 12)[0].tolist()
        got = out[rids[r]].tolist()
        n = next((i for i in range(12) if got[i] != solo[i]), 12)
        agree.append(n)
    ref_agree = [next((i for i in range(12) if out[rids[r]][i] != G["R2_tokens"][r][i]), 12) for r in range(10)]
    print(f"[serving R2] identical prefix vs own B=1 generate: {agree}; vs reference: {ref_agree}")
    assert sum(a >= 1 for a in ref_agree) >= 8


@pytest.mark.parametrize("regime", ["R0", "R1"])
def test_generate_with_prompt_lens_equals_single_request_reference_runs(regime):
    """generate(prompt_lens=...): a dense, right-padded batch of ragged prompts; every row must reproduce the unmodified
    reference's answer for that prompt alone (pads masked), for graph replay and eager decode."""
    model = build_model(TINY_CONFIG, make_state_dict(TINY_CONFIG, regime, seed=11))
    reqs = make_requests(TINY_CONFIG, 10, 2, 8, seed=21)
    lens = [int(i.numel()) for i, _ in reqs]
    S = max(lens)
    ids = torch.zeros(10, S, dtype=torch.int64)  # token 0 = pad id of the tiny config
    for r, (i, _) in enumerate(reqs):
        ids[r, : lens[r]] = i
    px = torch.stack([p for _, p in reqs])
    for graph in (True, False):
        toks = model.generate(ids.cuda(), px.cuda(), None, 12, prompt_lens=torch.tensor(lens), use_cuda_graph=graph)
        for r in range(10):
            assert toks[r].tolist() == G[f"{regime}_tokens"][r].tolist(), (graph, r)
    with pytest.raises(ValueError):
        model.generate(ids.cuda(), px.cuda(), None, 4, prompt_lens=torch.tensor([S + 1] * 10))


def test_batcher_input_validation():
    from paligemma_multimodal_system_b200.serving import ContinuousBatcher
    model = build_model(TINY_CONFIG, make_state_dict(TINY_CONFIG, "R1", seed=11))
    cb = ContinuousBatcher(model, num_slots=2, max_prompt_len=260, max_new_tokens=4)
    ids, px = make_requests(TINY_CONFIG, 1, 8, 8, seed=1)[0]
    with pytest.raises(ValueError):
        cb.submit(ids, px)  # 264 tokens > max_prompt_len
    with pytest.raises(ValueError):
        cb.submit(ids[:258], px, 5)  # budget above the batcher's
    bad = ids[:258].clone()
    bad[3] = 5  # 255 image tokens (the reference mis-scatters silently): refused at submit, not inside a prefill group
    with pytest.raises(ValueError):
        cb.submit(bad, px)
    with pytest.raises(ValueError):
        cb.submit(ids[:258], px[:, :100])
    oob = ids[:258].clone()
    oob[-1] = 5000
    with pytest.raises(IndexError):
        cb.submit(oob, px)
    rid = cb.submit(ids[:258], px)
    assert list(cb.run()) == [rid]  # the batcher is untouched by the refused requests


def test_long_generation_crosses_kv_page_boundaries():
    """140 new tokens from prompts of 258-263 tokens: the KV length passes the 64-token page boundaries at 320 and 384 in
    generate() (teacher-forced logits against the oracle at every step, diffuse regime) and in the batcher (free-running
    tokens against the oracle in the stable regime, slots refilled while others sit in their third page)."""
    from oracle import paligemma_oracle as O
    from paligemma_multimodal_system_b200.random_init import make_inputs
    from paligemma_multimodal_system_b200.serving import ContinuousBatcher
    T = 140
    sd = make_state_dict(TINY_CONFIG, "R2", seed=11)
    model = build_model(TINY_CONFIG, sd)
    inp = make_inputs(TINY_CONFIG, batch=2, prompt_len=5, seed=31)
    ref_t, ref_l = O.generate(sd, TINY_CONFIG, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], T, return_logits=True)
    _, logits = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), T,
                               return_logits=True, forced_tokens=ref_t)
    worst = max((stats(logits[r, t], ref_l[r, t])["rel"], r, t) for r in range(2) for t in range(T))
    cos = min(stats(logits[r, t], ref_l[r, t])["cos"] for r in range(2) for t in range(T))
    print(f"[long] teacher-forced R2, {T} steps: worst max-abs/absmax {worst[0]:.4f} at row {worst[1]} step {worst[2]}, min cosine {cos:.6f}")
    # tolerance of the diffuse tiny regime (3 % over 16 steps in test_model_parity_gpu.py); the maximum over 280 step-rows
    # sits a little higher (measured 2.8 %), hence 4 % here; the cosine gate is unchanged
    assert worst[0] <= 4e-2 and cos >= 0.999
    sd1 = make_state_dict(TINY_CONFIG, "R1", seed=11)
    model1 = build_model(TINY_CONFIG, sd1)
    reqs = make_requests(TINY_CONFIG, 5, 2, 7, seed=21)
    budgets = [T, 70, T, 9, 100]
    cb = ContinuousBatcher(model1, num_slots=2, max_prompt_len=264, max_new_tokens=T, steps_per_replay=8, stage=1)
    rids = [cb.submit(ids, px, m) for (ids, px), m in zip(reqs, budgets)]
    out = cb.run()
    for r, (ids, px) in enumerate(reqs):
        want = O.generate(sd1, TINY_CONFIG, ids[None], px[None], torch.ones(1, ids.numel(), dtype=torch.int64), budgets[r])[0].tolist()
        assert out[rids[r]].tolist() == want, r
