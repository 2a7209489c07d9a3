"""Read sharding across the GPUs of one box.

Reads are independent end to end, so the partition is the whole multi-GPU design: sort by
length, deal each read to the currently lightest rank (greedy LPT on sample count, which is
proportional to blocks and therefore to work).  No collective is ever needed on the data
path; results are gathered on the host and emitted in input order."""
from __future__ import annotations

import heapq
from typing import List

import numpy as np


def shard_reads(lengths, world: int) -> List[List[int]]:
    lengths = np.asarray(lengths, np.int64)
    shards: List[List[int]] = [[] for _ in range(world)]
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    for i in np.argsort(-lengths, kind="stable"):
        load, r = heapq.heappop(heap)
        shards[r].append(int(i))
        heapq.heappush(heap, (load + int(lengths[i]), r))
    return shards
