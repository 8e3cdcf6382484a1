"""N>1 host path on the CPU: two gloo ranks shard the request list with no data-path collective and agree on the
max-over-ranks timing / aggregate throughput that bench.py reports."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opsagent_b200.dp import aggregate_throughput, allreduce_max, shard_requests


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = list(shard_requests(1024, rank, world))
    seconds = 2.0 + rank            # rank 1 is the slow replica
    mx = allreduce_max(seconds, dist)
    thr = aggregate_throughput(len(mine) * 256, seconds, dist)
    counts = torch.tensor([len(mine)]); dist.all_reduce(counts)
    q.put((rank, mine[0], mine[-1], mx, thr, int(counts.item())))
    dist.barrier(); dist.destroy_process_group()


def test_two_rank_dp_sharding_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert res[0][1:3] == (0, 511) and res[1][1:3] == (512, 1023)
    for r in res:
        assert r[3] == 3.0                                   # slowest rank's time
        assert abs(r[4] - 2 * 512 * 256 / 3.0) < 1e-6        # whole-job units / max time
        assert r[5] == 1024
