#!/usr/bin/env python3
"""Print the clock trace written by the single-pass kernel under RB200_OP_TRACE=1 (CTA 0, first 32 items).
Events per (item, warp): 0 loop top, 1 first pair ready, 2 pair loop exit, 3 before barrier A, 4 after A, 5 before barrier B,
6 after B, 7 Doppler column done, 8 pairs taken, 9 cycles in mbarrier waits, 10/11 half A/B refill issued, 12 cycles in fetch."""
import sys
import numpy as np

tr = np.fromfile(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/op_trace.bin", dtype=np.uint64).reshape(32, 12, 16).astype(np.int64)
t0 = tr[0, :, 0].min()
for k in range(int(sys.argv[2]) if len(sys.argv) > 2 else 8):
    e = tr[k]
    if e[:, 0].max() == 0:
        break
    base = e[:, 0].min()
    print("item %d: starts at %d clk (item period %d)" % (k, base - t0, base - tr[k - 1, :, 0].min() if k else 0))
    for w in range(12):
        rel = lambda ev: (e[w, ev] - base) if e[w, ev] else -1
        print("  w%02d top %6d first-pair %6d loop-exit %6d A-in %6d A-out %6d B-in %6d B-out %6d dop-done %6d | pairs %d mbar-wait %6d fetch %6d refillA %6d refillB %6d"
              % (w, rel(0), rel(1), rel(2), rel(3), rel(4), rel(5), rel(6), rel(7), e[w, 8], e[w, 9], e[w, 12], rel(10), rel(11)))
