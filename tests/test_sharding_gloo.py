"""world_size-2 CPU (gloo) check of the multi-GPU plumbing: frame ranges are disjoint and complete across ranks,
and the timing reduction bench.py uses (max over ranks) behaves.  No data-path collective exists (frames are independent)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from zstdsharp_b200.sharding import my_shard, shard_bounds_native
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    weights = (np.arange(1000) % 97 + 22).tolist()
    lo, hi = my_shard(weights, rank, world)
    assert shard_bounds_native(weights, world)[rank] == (lo, hi)       # the in-library scatter (ZSTDB200_*BatchMulti) owns the same range
    mine = torch.zeros(1000, dtype=torch.int32)
    mine[lo:hi] = 1
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)           # test-only collective: every frame owned exactly once
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, lo, hi, bool((mine == 1).all()), float(t.item())))
    dist.destroy_process_group()


def test_two_rank_partition():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 1000
    assert all(r[3] for r in res) and all(r[4] == 2.0 for r in res)
