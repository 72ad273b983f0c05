"""world_size-2 gloo test of the multi-GPU plumbing: CPI block sharding and the all-gather of the sparse
detection lists (the only inter-rank exchange of the path; NCCL on the GPUs, gloo here)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from radar_signal_process_b200 import DET_DTYPE
from radar_signal_process_b200 import distributed as rdist


def test_shard_range_partitions_exactly():
    for n, w in ((64, 8), (10, 4), (3, 4), (0, 2), (7, 1)):
        spans = [rdist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make(rank, n):
    d = np.zeros(n, dtype=DET_DTYPE)
    d["cpi"] = np.arange(n) % 3
    d["lane"] = rank
    d["v"] = np.arange(n) % 64
    d["r"] = 1000 * rank + np.arange(n)
    d["kind"] = 2
    d["amp"] = np.arange(n, dtype=np.float32) + rank
    return d


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_local = 5 + 7 * rank                       # ragged: 5 and 12 detections; rank 1 list may be empty in 2nd round
        allrec, counts = rdist.gather_detections(_make(rank, n_local), capacity=32, cpi_offset=10 * rank)
        empty, counts2 = rdist.gather_detections(np.zeros(0, dtype=DET_DTYPE), capacity=4)
        over, counts3 = rdist.gather_detections(_make(rank, 9), capacity=4)      # truncated to capacity
        q.put((rank, allrec.tobytes(), counts.tolist(), len(empty), counts2.tolist(), len(over), counts3.tolist()))
    finally:
        dist.destroy_process_group()


def test_gather_detections_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.concatenate([_make(0, 5), _make(1, 12)])
    want["cpi"][5:] += 10
    for rank, blob, counts, n_empty, counts2, n_over, counts3 in res:
        got = np.frombuffer(blob, dtype=DET_DTYPE)
        assert counts == [5, 12]
        assert np.array_equal(got, want)              # every rank holds the same rank-ordered list, global CPI indices
        assert n_empty == 0 and counts2 == [0, 0]
        assert n_over == 8 and counts3 == [4, 4]
