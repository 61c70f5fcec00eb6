"""N>1 path on CPU: contiguous batch sharding and the single counter all-reduce,
exercised with world_size-2 gloo processes."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crog_b200.engine import GraspEvaluator, shard_range


def test_shard_range_partitions_without_duplicates():
    for n in (0, 1, 7, 64, 65, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    ev = GraspEvaluator(model=None, device=torch.device("cpu"))
    # stand-in for the per-rank device counters: sample i is "correct@1" if i % 3 == 0, "correct@5" if i % 2 == 0
    for i in range(lo, hi):
        ev.counters += torch.tensor([int(i % 3 == 0), 1, int(i % 2 == 0), 1])
    out = ev.reduce().tolist()
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_counter_allreduce_two_ranks():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    n = 37
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    [p.start() for p in procs]
    out = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert out == [len(range(0, n, 3)), n, len(range(0, n, 2)), n]
