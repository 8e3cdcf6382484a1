"""dp — host-side helpers of the one-process-per-GPU replica mode that bench.py measures (BASELINE configs[2]: one engine per GPU,
requests are independent, NO data-path collective): which request ids a rank owns, and the max-over-ranks timing / whole-job throughput
the JSON line reports.  torch.distributed is used only for that barrier/max.  Serving the same replicas behind ONE endpoint — sticky
routing of live conversations, per-replica admission — is opsagent_b200/router.py."""
from __future__ import annotations


def shard_requests(n_requests: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of request indices for this replica (sticky per conversation: a request id always
    maps to the same replica, so its prefix KV stays on one GPU)."""
    base, rem = divmod(n_requests, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def replica_of(request_id: int, n_requests: int, world: int) -> int:
    base, rem = divmod(n_requests, world)
    edge = rem * (base + 1)
    return request_id // (base + 1) if request_id < edge else rem + (request_id - edge) // max(base, 1)


def allreduce_max(value: float, dist=None, device=None) -> float:
    """max over ranks of a host scalar (device timing of the slowest replica)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank: int, seconds: float, dist=None, device=None) -> float:
    """whole-job throughput = units all ranks processed / slowest rank's time"""
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    return world * units_per_rank / allreduce_max(seconds, dist, device)
