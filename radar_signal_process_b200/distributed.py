"""Multi-GPU plumbing: CPIs are block-partitioned over ranks (one process per GPU, no inter-GPU
traffic on the hot path); only the sparse detection lists are exchanged, with NCCL over NVLink
(``torch.distributed`` all_gather; gloo on CPU for the tests).  SURVEY.md section 8(e).
"""
import numpy as np
import torch
import torch.distributed as dist

from ._binding import DET_DTYPE


def shard_range(n_total, rank, world):
    """Contiguous block of CPI indices [lo, hi) owned by ``rank`` (blocks differ by at most one)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_detections(dets, capacity, cpi_offset=0, device=None, group=None):
    """All-gather per-rank detection lists.  ``dets``: structured array (DET_DTYPE) of this rank with
    rank-local CPI indices; ``cpi_offset`` is added so the gathered list carries global CPI indices.
    Two collectives: the counts (one int32 per rank), then fixed-capacity padded 16-byte records.
    Returns (concatenated structured array ordered by rank, counts per rank).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dets = np.asarray(dets, dtype=DET_DTYPE)
    n = min(len(dets), int(capacity))
    local = dets[:n].copy()
    local["cpi"] += np.uint32(cpi_offset)
    if world == 1:
        return local, np.array([n], dtype=np.int64)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    cnts = torch.zeros(world, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    buf = np.zeros(int(capacity), dtype=DET_DTYPE)
    buf[:n] = local
    t = torch.from_numpy(buf.view(np.uint8).reshape(-1)).to(dev)
    out = torch.zeros(world * t.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, t, group=group)
    counts = cnts.cpu().numpy().astype(np.int64)
    allrec = out.cpu().numpy().view(DET_DTYPE).reshape(world, int(capacity))
    return np.concatenate([allrec[r, :counts[r]] for r in range(world)]), counts
