"""Multi-GPU plumbing: CPIs are block-partitioned over ranks (one process per GPU, no inter-GPU
traffic on the hot path); only the sparse detection lists are exchanged, with NCCL over NVLink
(``torch.distributed`` all_gather; gloo on CPU for the tests).  SURVEY.md section 8(e).

Two forms of the gather:
  * ``gather_detections_device`` -- the production form: the records never leave the device.  The lists are read where
    the chain left them (``rb200_chain_dets_device``), the per-rank counts are exchanged (2 x int32 per rank), and every
    rank contributes exactly ``max(count)`` records to one ``all_gather`` -- no padding to ``max_det``, no host bounce.
  * ``gather_detections`` -- host structured arrays in, host arrays out (used by the CPU/gloo tests and by callers that
    already hold the lists on the host).
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from ._binding import DET_DTYPE

_REC = 16   # bytes per rb200_det


def shard_range(n_total, rank, world):
    """Contiguous block of CPI indices [lo, hi) owned by ``rank`` (blocks differ by at most one)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _exchange(local, n, cpi_offset, group):
    """local: int32 tensor [cap, 4] (16-byte records, cpi in column 0) on the collective's device, n valid rows."""
    world = dist.get_world_size(group)
    dev = local.device
    cnt = torch.tensor([n], dtype=torch.int32, device=dev)
    cnts = torch.empty(world, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    counts = cnts.cpu().numpy().astype(np.int64)
    m = int(counts.max())
    if m == 0:
        return torch.empty((0, 4), dtype=torch.int32, device=dev), counts
    if local.shape[0] < m:                                     # every rank contributes exactly max(count) records
        grown = torch.zeros((m, 4), dtype=torch.int32, device=dev)
        grown[:n] = local[:n]
        local = grown
    send = local[:m]
    if cpi_offset:
        send = send.clone()
        send[:n, 0] += int(cpi_offset)
    out = torch.empty((world * m, 4), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    keep = torch.cat([out[r * m:r * m + int(counts[r])] for r in range(world)])
    return keep, counts


class DeviceRecords:
    """A torch int32 [n, 4] view over ``n`` 16-byte detection records that live in device memory owned by the library."""

    def __init__(self, ptr, n, device):
        self.n = int(n)
        self.tensor = torch.empty((0, 4), dtype=torch.int32, device=device)
        if self.n:
            nbytes = self.n * _REC

            class _Iface:                                      # __cuda_array_interface__: zero-copy wrap of the pointer
                __cuda_array_interface__ = {"shape": (self.n, 4), "typestr": "<i4", "data": (int(ptr), False), "version": 3,
                                            "strides": None}
            self._keep = _Iface()
            self.tensor = torch.as_tensor(self._keep, device=device)
            assert self.tensor.numel() * 4 == nbytes


def gather_detections_device(ctx, cpi_offset=0, device=None, group=None, kind="2d"):
    """All-gather the detection list of the last chain call of ``ctx`` without leaving the device.
    Returns (int32 tensor [total, 4] on the device -- view it as DET_DTYPE after ``.cpu().numpy()`` --, counts per rank)."""
    (p2, n2), (pv, nv) = ctx.chain_dets_device(allow_overflow=True)
    ptr, n = (p2, n2) if kind == "2d" else (pv, nv)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    rec = DeviceRecords(ptr, n, dev).tensor
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = rec.clone()
        if cpi_offset and n:
            out[:, 0] += int(cpi_offset)
        return out, np.array([n], dtype=np.int64)
    return _exchange(rec, n, cpi_offset, group)


def records_to_numpy(t):
    """int32 [n, 4] tensor (any device) -> structured DET_DTYPE array."""
    return np.ascontiguousarray(t.cpu().numpy()).view(np.uint8).reshape(-1).view(DET_DTYPE)


def gather_detections(dets, capacity=None, cpi_offset=0, device=None, group=None):
    """All-gather per-rank detection lists held on the host.  ``dets``: structured array (DET_DTYPE) of this rank with
    rank-local CPI indices; ``cpi_offset`` is added so the gathered list carries global CPI indices.  Two collectives:
    the counts (one int32 per rank), then exactly max(count) 16-byte records per rank.
    Returns (concatenated structured array ordered by rank, counts per rank)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dets = np.ascontiguousarray(np.asarray(dets, dtype=DET_DTYPE))
    n = len(dets) if capacity is None else min(len(dets), int(capacity))
    if world == 1:
        local = dets[:n].copy()
        local["cpi"] += np.uint32(cpi_offset)
        return local, np.array([n], dtype=np.int64)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    rec = torch.from_numpy(dets[:n].view(np.uint8).reshape(-1).view(np.int32).reshape(-1, 4).copy()).to(dev)
    keep, counts = _exchange(rec, n, cpi_offset, group)
    return records_to_numpy(keep), counts
