#!/usr/bin/env python3
"""Run the same device-resident batch repeatedly and report any run-to-run difference in RDM bits or detections."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radar_signal_process_b200 as rsp
from radar_signal_process_b200 import waveforms, workload

P, R, C, B = 64, 4096, 16, 16
raw = workload.synth_batch(B, distinct=8)
dev = torch.device("cuda", 0)
raw_d = torch.from_numpy(raw).to(dev)
rdm_d = torch.zeros((B, C, P, R), dtype=torch.float32, device=dev)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ctx = rsp.Context(0, n_prt=P, n_range=R, n_lanes=C, max_cpi=B, max_det=1 << 20)
ctx.set_waveform(waveforms.segments_single(R, waveforms.REF_DDC))
ctx.set_cfar(5, 7, 5.0, 0, 5, 7, 5.0, 0, 0, 1)


def key(d):
    return np.sort(d, order=["cpi", "lane", "v", "r", "kind"])


ref_rdm = ref_d = None
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    rdm_d.zero_()
    ctx.chain_enqueue(raw_d.data_ptr(), B, rdm_d.data_ptr(), stream.cuda_stream)
    dets, n = ctx.chain_fetch()
    stream.synchronize()
    rdm = rdm_d.cpu().numpy()
    d = key(dets)
    if ref_rdm is None:
        ref_rdm, ref_d = rdm, d
        print("run 0: n_det", n)
        continue
    nd = int((rdm != ref_rdm).sum())
    same_d = len(d) == len(ref_d) and all(np.array_equal(d[f], ref_d[f]) for f in d.dtype.names)
    if nd or not same_d:
        bad += 1
        idx = np.argwhere(rdm != ref_rdm)
        print("run %d: %d RDM cells differ (first %s), n_det %d vs %d, dets equal %s" % (it, nd, idx[:3].tolist(), n, len(ref_d), same_d))
        if nd:
            i = tuple(idx[0])
            print("   values", rdm[i], ref_rdm[i])
print("runs with differences:", bad)
ctx.close()
