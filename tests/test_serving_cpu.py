"""Host-side logic of serving.py (no GPU): slot scheduling and the reference loop's stop rule (inference.py:71-74), and the
oracle pinned on the ragged request stream against vectors of the unmodified reference (tests/golden/make_serving_golden.py)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import paligemma_oracle as O  # noqa: E402
from paligemma_multimodal_system_b200.random_init import TINY_CONFIG, make_requests, make_state_dict  # noqa: E402
from paligemma_multimodal_system_b200.serving import Request, SlotScheduler  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "serving_reference.npz"))


def _req(rid, max_new):
    return Request(rid, torch.zeros(4, dtype=torch.int64), torch.zeros(3, 2, 2), max_new)


def _admit(s):
    """prefill everything the scheduler wants prefilled (first token 1), then arm: the stage = 0 behaviour."""
    for r in s.plan_prefill():
        s.prefilled(r, 1)
    return [(slot, r.rid) for slot, r in s.plan_arming()]


def test_scheduler_fifo_lowest_slot_and_refill():
    s = SlotScheduler(2)
    for i in range(5):
        s.submit(_req(i, 3))
    assert _admit(s) == [(0, 0), (1, 1)]
    assert _admit(s) == []  # no free page set
    assert not s.consume(0, [7])
    assert s.consume(1, [2, 3, 4, 5])  # budget 3 (first token came from the prefill): the replay's overshoot is dropped
    assert s.finished[1] == [1, 2, 3]
    assert _admit(s) == [(1, 2)]
    assert s.consume(0, [9])
    assert s.finished[0] == [1, 7, 9]
    assert not s.idle()
    assert _admit(s) == [(0, 3)]
    s.consume(0, [1, 1]); s.consume(1, [1, 1])
    assert _admit(s) == [(0, 4)]
    s.consume(0, [1, 1])
    assert s.idle() and sorted(s.finished) == [0, 1, 2, 3, 4]
    assert sorted(s.free_sets) == [0, 1] and sorted(s.free_slots) == [0, 1]


def test_scheduler_eos_is_appended_then_stops():
    s = SlotScheduler(1, eos_token_id=1)
    s.submit(_req(0, 10))
    (r,) = s.plan_prefill()
    s.prefilled(r, 5)
    s.plan_arming()
    assert not s.consume(0, [6])
    assert s.consume(0, [7, 1, 9, 9])
    assert s.finished[0] == [5, 6, 7, 1]
    s.submit(_req(1, 10))
    (r,) = s.plan_prefill()
    s.prefilled(r, 1)  # EOS as the very first token: the request never takes a slot, its page set is free again
    assert s.finished[1] == [1] and s.plan_arming() == [] and s.idle() and s.free_sets == [0]
    s.submit(_req(2, 1))
    (r,) = s.plan_prefill()
    s.prefilled(r, 7)  # budget of one token: done at the prefill
    assert s.finished[2] == [7] and s.idle()


def test_scheduler_min_admit_waits_for_a_group():
    s = SlotScheduler(4, min_admit=2)
    for i in range(7):
        s.submit(_req(i, 3))
    assert len(_admit(s)) == 4
    s.consume(0, [1, 1])
    assert _admit(s) == []  # one free set < min_admit while others are busy
    s.consume(2, [1, 1])
    assert [slot for slot, _ in _admit(s)] == [0, 2]
    s.consume(1, [1, 1])
    # only one request is left in the queue: it does not wait for a second one
    assert _admit(s) == [(1, 6)]
    for slot in list(s.active):
        s.consume(slot, [1, 1])
    assert s.idle()
    s2 = SlotScheduler(4, min_admit=3)
    s2.submit(_req(0, 2))
    assert len(_admit(s2)) == 1  # nothing is running: never wait


def test_scheduler_stages_ahead_and_hands_over_page_sets():
    """2 slots + 2 extra page sets: requests 2, 3 are prefilled while 0, 1 decode, and take a slot over the moment it frees."""
    s = SlotScheduler(2, min_admit=2, stage=2)
    for i in range(6):
        s.submit(_req(i, 3))
    reqs = s.plan_prefill()
    assert [(r.rid, r.page_set) for r in reqs] == [(0, 0), (1, 1), (2, 2), (3, 3)]
    for r in reqs:
        s.prefilled(r, 1)
    assert [(slot, r.rid) for slot, r in s.plan_arming()] == [(0, 0), (1, 1)]
    assert [r.rid for r in s.staged] == [2, 3] and s.plan_prefill() == []  # no free set
    assert s.consume(1, [1, 1])                      # request 1 done: slot 1 and set 1 free
    assert s.plan_prefill() == []                    # one free set < min_admit
    assert [(slot, r.rid, r.page_set) for slot, r in s.plan_arming()] == [(1, 2, 2)]  # staged request 2 takes slot 1 at once
    assert s.consume(0, [1, 1])
    reqs = s.plan_prefill()                          # sets 0 and 1 are free: the next group is prefilled ahead
    assert [(r.rid, r.page_set) for r in reqs] == [(4, 0), (5, 1)]
    for r in reqs:
        s.prefilled(r, 1)
    assert [(slot, r.rid) for slot, r in s.plan_arming()] == [(0, 3)]
    while not s.idle():
        for slot in list(s.active):
            s.consume(slot, [1, 1])
        s.plan_arming()
    assert sorted(s.finished) == list(range(6)) and sorted(s.free_sets) == [0, 1, 2, 3]


def test_scheduler_rejects_bad_sizes():
    with pytest.raises(ValueError):
        SlotScheduler(0)


@pytest.mark.parametrize("regime", ["R0", "R1", "R2"])
def test_oracle_matches_reference_on_ragged_requests(regime):
    sd = make_state_dict(TINY_CONFIG, regime, seed=11)
    reqs = make_requests(TINY_CONFIG, 10, 2, 8, seed=21)
    assert [int(i.numel()) for i, _ in reqs] == G["prompt_lens"].tolist()
    for r in (0, 1, 4, 9):
        ids, px = reqs[r]
        toks, logits = O.generate(sd, TINY_CONFIG, ids[None], px[None], torch.ones(1, ids.numel(), dtype=torch.int64), 12,
                                  return_logits=True)
        assert toks[0].tolist() == G[f"{regime}_tokens"][r].tolist()
        ref = torch.as_tensor(G[f"{regime}_prefill_logits"][r])
        assert (logits[0, 0] - ref).abs().max().item() <= 2e-5 * max(ref.abs().max().item(), 1.0)
