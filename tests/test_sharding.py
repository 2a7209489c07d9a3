"""Multi-GPU host logic on CPU: reads shard embarrassingly across ranks with no data-path
collective; world_size-2 gloo covers the rank plumbing bench.py uses
(barrier + MAX-reduce of the step time, weak scaling bookkeeping)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flappie_b200.shard import shard_reads


def test_shard_reads_balances_blocks():
    rng = np.random.default_rng(0)
    lens = np.exp(rng.uniform(np.log(1000), np.log(50000), size=257)).astype(np.int64)
    for world in (1, 2, 4, 8):
        shards = shard_reads(lens, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(lens)))                     # a partition: every read exactly once
        loads = [int(lens[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= lens.max()              # greedy LPT bound
    assert shard_reads(np.array([], np.int64), 4) == [[], [], [], []]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, lens, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_reads(lens, world)[rank]
    # each rank would basecall only its shard; the only communication is timing bookkeeping
    t = torch.tensor([float(10 + rank)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(lens[mine].sum())], dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    out_q.put((rank, sorted(mine), float(t.item()), float(n.item())))
    dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    lens = np.arange(1000, 1000 + 37 * 100, 100, dtype=np.int64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert sorted(res[0][1] + res[1][1]) == list(range(len(lens)))
    assert res[0][2] == res[1][2] == 11.0                        # MAX over ranks
    assert res[0][3] == res[1][3] == float(lens.sum())           # every sample counted once
