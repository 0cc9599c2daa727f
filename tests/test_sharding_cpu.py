"""world_size-2 gloo test (CPU) of the batch-sharded replica plumbing: shards cover the batch exactly once, the oracle
run on each shard equals the oracle run on the whole batch (requests are independent: no data-path collective), token
gather and MAX timing reduction work."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_bounds_cover_batch():
    from paligemma_multimodal_system_b200.sharding import shard_bounds
    for B in (1, 7, 8, 64, 65):
        for N in (1, 2, 4, 8):
            spans = [shard_bounds(B, N, r) for r in range(N)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(N - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def test_shard_stream_round_robin():
    from paligemma_multimodal_system_b200.sharding import shard_stream
    for n in (0, 1, 7, 64):
        for N in (1, 2, 8):
            parts = [shard_stream(n, N, r) for r in range(N)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_stream(8, 2, -1)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import paligemma_oracle as O
    from paligemma_multimodal_system_b200.random_init import TINY_CONFIG, make_inputs, make_state_dict
    from paligemma_multimodal_system_b200.sharding import gather_tokens, max_over_ranks, shard_requests
    sd = make_state_dict(TINY_CONFIG, "R2", seed=3)  # identical replica on every rank (seeded)
    batch = make_inputs(TINY_CONFIG, batch=3, prompt_len=5, seed=9)
    mine = shard_requests(batch, world, rank)
    toks = O.generate(sd, TINY_CONFIG, mine["input_ids"], mine["pixel_values"], mine["attention_mask"], 4)
    allt = gather_tokens(toks, 3)
    tmax = max_over_ranks([float(rank + 1), 5.0 - rank], "cpu")
    # ragged request stream (serving): round-robin over the replicas, each request served on its own (as the reference does)
    from paligemma_multimodal_system_b200.random_init import make_requests
    from paligemma_multimodal_system_b200.sharding import gather_stream_results, shard_stream
    sd1 = make_state_dict(TINY_CONFIG, "R1", seed=11)
    reqs = make_requests(TINY_CONFIG, 5, 2, 8, seed=21)
    local = {}
    for i in shard_stream(len(reqs), world, rank):
        ids, px = reqs[i]
        local[i] = O.generate(sd1, TINY_CONFIG, ids[None], px[None], torch.ones(1, ids.numel(), dtype=torch.int64), 3)[0].tolist()
    stream = gather_stream_results(local)
    if rank == 0:
        ref = O.generate(sd, TINY_CONFIG, batch["input_ids"], batch["pixel_values"], batch["attention_mask"], 4)
        q.put((allt.tolist(), ref.tolist(), tmax, stream))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replicas_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, ref, tmax, stream = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got == ref
    assert tmax == [2.0, 5.0]
    import numpy as np
    G = np.load(os.path.join(ROOT, "tests", "golden", "serving_reference.npz"))
    assert sorted(stream) == [0, 1, 2, 3, 4]
    for i in range(5):  # every request, whichever rank served it, equals the unmodified reference's answer for it
        assert stream[i] == G["R1_tokens"][i][:3].tolist()
