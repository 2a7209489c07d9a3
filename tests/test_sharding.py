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


def _plan(lib, T, max_clusters, slots_max, can_stream=1):
    import ctypes
    from ctypes import POINTER, c_int, c_int32, c_int64
    L = lib.lib
    L.ffb_plan_schedule.restype = c_int64
    L.ffb_plan_schedule.argtypes = [POINTER(c_int64), c_int64, c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32),
                                    POINTER(c_int32), POINTER(c_int), POINTER(c_int)]
    T = np.ascontiguousarray(T, np.int64)
    n = T.shape[0]
    groups = (n + 15) // 16
    order = np.full(max(groups * 16, 1), -2, np.int32)
    slot_off = np.zeros(max_clusters * slots_max + 1, np.int32)
    slot_list = np.full(max(groups, 1), -2, np.int32)
    ncl, G = c_int(0), c_int(0)
    r = L.ffb_plan_schedule(T.ctypes.data_as(POINTER(c_int64)), n, max_clusters, slots_max, can_stream,
                            order.ctypes.data_as(POINTER(c_int32)), slot_off.ctypes.data_as(POINTER(c_int32)),
                            slot_list.ctypes.data_as(POINTER(c_int32)), ctypes.byref(ncl), ctypes.byref(G))
    assert r == groups
    return order[:groups * 16], slot_off[:ncl.value * G.value + 1], slot_list[:groups], ncl.value, G.value


def test_recurrent_group_schedule(lib):
    """The host schedule of the tensor recurrent kernel (csrc/api.cu:plan_groups, no device needed): every group runs
    exactly once, slots are balanced (LPT bound), a batch that fits one wave gets one group per slot."""
    # configs[1]: 1024 equal reads, 15 co-resident clusters of 5 slots -> 13 clusters x 5 slots, one group each
    order, so, sl, ncl, G = _plan(lib, np.full(1024, 1895), 15, 5)
    assert (ncl, G) == (13, 5) and sorted(sl.tolist()) == list(range(64))
    assert np.all(np.diff(so) <= 1) and sorted(order.tolist()) == list(range(1024))
    # configs[2]: 4096 reads -> more groups than slots: every slot of all 15 clusters busy, 3-4 groups each
    order, so, sl, ncl, G = _plan(lib, np.full(4096, 1895), 15, 5)
    assert (ncl, G) == (15, 5) and sorted(sl.tolist()) == list(range(256))
    assert set(np.diff(so).tolist()) <= {3, 4}
    # ... and with SIX slots per cluster available (S = 256): six slots x all 15 clusters = 90 slots finish the 256 groups in
    # three rounds instead of four, which outweighs the slower step (profiles/r02_slots_ab.txt); the one-wave batch stays
    # on five slots
    order, so, sl, ncl, G = _plan(lib, np.full(4096, 1895), 15, 6)
    assert (ncl, G) == (15, 6) and sorted(sl.tolist()) == list(range(256)) and set(np.diff(so).tolist()) <= {2, 3}
    assert _plan(lib, np.full(1024, 1895), 15, 6)[3:] == (13, 5)
    # no streamed GEMM (K > 256): no SMs held back
    assert _plan(lib, np.full(4096, 758), 12, 4, can_stream=0)[3] == 12
    # ragged batch (configs[3] lengths): a partition, groups sorted by length, loads within one longest group of each other
    rng = np.random.default_rng(2)
    T = np.exp(rng.uniform(np.log(200), np.log(5000), size=4000)).astype(np.int64)
    order, so, sl, ncl, G = _plan(lib, T, 15, 5)
    assert sorted(order[order >= 0].tolist()) == list(range(4000)) and np.all(order[4000:] == -1)
    Tg = T[order[::16]]                                       # longest read of every group
    assert np.all(np.diff(Tg) <= 0) and sorted(sl.tolist()) == list(range(250))
    loads = np.array([Tg[sl[so[k]:so[k + 1]]].sum() for k in range(ncl * G)])
    assert loads.max() - loads.min() <= Tg.max()
    # a batch whose longest read outlasts the average slot (configs[3]: 1 k - 50 k): the makespan is that read whatever the
    # slot count, so the planner takes FEWER slots per cluster -- a shorter step chain -- as long as the makespan stays there
    T = np.exp(rng.uniform(np.log(395), np.log(24895), size=1024)).astype(np.int64)
    order, so, sl, ncl, G = _plan(lib, T, 15, 6)
    Tg = T[order[::16]]
    loads = np.array([Tg[sl[so[k]:so[k + 1]]].sum() for k in range(ncl * G)])
    assert G < 5 and loads.max() == Tg.max()
    # empty batch and a batch smaller than one group
    assert _plan(lib, np.zeros(0, np.int64), 15, 5)[3] == 0
    order, so, sl, ncl, G = _plan(lib, np.array([700, 10, 0]), 15, 5)
    assert (ncl, G) == (1, 1) and order[:3].tolist() == [0, 1, 2] and np.all(order[3:] == -1)


def test_c_lpt_deal_matches_the_python_rule():
    """ffb_deal_lpt (flappie_b200/host/ffb_shard.c, what `flappie --devices` uses per window) against shard_reads: same
    loads per device, capacity respected, unreadable reads dealt to nobody, deterministic."""
    import ctypes
    import os
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = ctypes.CDLL(os.path.join(here, "flappie_b200", "host", "libffb_host.so"))
    host.ffb_deal_lpt.restype = ctypes.c_int
    host.ffb_deal_lpt.argtypes = [ctypes.POINTER(ctypes.c_long), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    rng = np.random.default_rng(3)
    for ndev, n in ((8, 8 * 1024), (2, 301), (3, 7), (4, 0)):
        lens = np.exp(rng.uniform(np.log(1000), np.log(50000), n)).astype(np.int64)
        if n > 10:
            lens[[3, 9]] = [0, -2]                        # unreadable / fast5 without libhdf5
        cl = (ctypes.c_long * max(n, 1))(*lens.tolist())
        dev = (ctypes.c_int * max(n, 1))()
        cap = (n + ndev - 1) // ndev + 2
        dealt = host.ffb_deal_lpt(cl, n, ndev, cap, dev)
        d = np.array(dev[:n])
        assert dealt == int(np.sum(lens > 0)) and np.all(d[lens <= 0] == -1) and np.all(d[lens > 0] >= 0)
        loads = np.array([lens[d == k].sum() for k in range(ndev)])
        counts = np.array([np.sum(d == k) for k in range(ndev)])
        assert counts.max(initial=0) <= cap
        if n > 10:
            assert loads.max() - loads.min() <= lens.max()           # LPT: within one (longest) read of each other
            ref = shard_reads(np.where(lens > 0, lens, 0), ndev)
            ref_loads = sorted(int(lens[s][lens[s] > 0].sum()) for s in ref)
            assert abs(max(ref_loads) - loads.max()) <= lens.max()
        dev2 = (ctypes.c_int * max(n, 1))()
        host.ffb_deal_lpt(cl, n, ndev, cap, dev2)
        assert list(dev2[:n]) == list(dev[:n])
    assert host.ffb_deal_lpt(None, 1, 1, 1, None) == -1
