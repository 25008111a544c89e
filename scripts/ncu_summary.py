#!/usr/bin/env python
"""Summarise an ncu --set full report: one line of key metrics per kernel launch.
usage: scripts/ncu_summary.py report.ncu-rep [kernel-substring]"""
import csv, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active", "dmma%"),
    ("sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_cyc%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "lg_throttle", "mio_throttle", "math_pipe_throttle",
               "not_selected", "dispatch_stall", "branch_resolving", "no_instruction", "membar", "sleeping", "drain"]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    for row in rows[2:]:
        if sub not in row[ki]:
            continue
        print("##", row[ki][:70])
        parts = []
        for k, nm in KEYS:
            if k in h:
                i = h.index(k)
                parts.append("%s=%s%s" % (nm, row[i], units[i] if units[i] not in ("%", "") and nm not in ("regs",) else ""))
        print("  ", "  ".join(parts))
        st = []
        for s in STALL_NAMES:
            k = STALLS % s
            if k in h:
                v = float(row[h.index(k)] or 0)
                if v >= 0.2:
                    st.append((v, s))
        print("   stalls/issue:", ", ".join("%s %.1f" % (s, v) for v, s in sorted(st, reverse=True)))


if __name__ == "__main__":
    main()
