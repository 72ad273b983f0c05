#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of per-kernel numbers quoted in profiles/*.md.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [max_kernels]"""
import csv
import subprocess
import sys

WANT = [
    ("duration_us", "gpu__time_duration.sum"),
    ("dram_read_MB", "dram__bytes_read.sum"),
    ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("waves", "launch__waves_per_multiprocessor"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("pipe_fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("pipe_alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("pipe_lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("pipe_xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
]


def main():
    rep = sys.argv[1]
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:2 + limit]:
        print("== %s" % r[ki][:90])
        for name, key in WANT:
            if key in hdr:
                i = hdr.index(key)
                print("   %-22s %s %s" % (name, r[i], units[i]))
        st = []
        for i in stall_cols:
            try:
                st.append((float(r[i].replace(",", "")), hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
        st.sort(reverse=True)
        print("   stalled warps per issue-active cycle: " + ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))


if __name__ == "__main__":
    main()
