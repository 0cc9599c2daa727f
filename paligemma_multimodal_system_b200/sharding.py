"""Batch sharding of independent requests over data-parallel replicas (one process per GPU, no data-path collective).

Rank r of N serves the contiguous rows [r*B/N, (r+1)*B/N) of a global request batch with its own full weight replica
and KV cache; only timings / token ids are gathered for reporting (NCCL on GPUs, gloo in the CPU tests)."""
import torch


def shard_bounds(global_batch: int, world_size: int, rank: int):
    """Contiguous, balanced split: the first (global_batch % world_size) ranks get one extra row."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    base, extra = divmod(global_batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_requests(batch: dict, world_size: int, rank: int) -> dict:
    lo, hi = shard_bounds(next(iter(batch.values())).shape[0], world_size, rank)
    return {k: v[lo:hi] for k, v in batch.items()}


def shard_stream(n_requests: int, world_size: int, rank: int):
    """Request stream over replicas (serving.ContinuousBatcher per GPU): round-robin, so that ragged prompt / answer
    lengths in arrival order spread evenly.  Returns the indices rank `rank` serves."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    return list(range(rank, n_requests, world_size))


def gather_stream_results(local: dict, group=None) -> dict:
    """{request index: token list} of every rank merged on all ranks (reporting only; off the timed path)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, {int(k): [int(t) for t in v] for k, v in local.items()}, group=group)
    out = {}
    for part in parts:
        assert not (set(part) & set(out)), "a request was served by two ranks"
        out.update(part)
    return out


def gather_tokens(local_tokens: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """All-gathers the generated token ids [b_local, T] into [global_batch, T] (reporting only; off the timed path)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_tokens
    world = dist.get_world_size(group)
    T = local_tokens.shape[1]
    sizes = [shard_bounds(global_batch, world, r) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(pad, T, dtype=local_tokens.dtype, device=local_tokens.device)
    buf[: local_tokens.shape[0]] = local_tokens
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def max_over_ranks(values, device, group=None):
    """Element-wise MAX of a list of floats over all ranks (how multi-GPU timings are reported)."""
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.tolist()
