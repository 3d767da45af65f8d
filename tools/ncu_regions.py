#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: headline metrics + stall samples split at the mbarrier waits.
usage: python tools/ncu_regions.py gpurun_out/prof_X.source.csv [gpurun_out/prof_X.raw.csv]"""
import csv
import sys

csv.field_size_limit(10 ** 9)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = rows[1]
    src, ex, si = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    data = [(k, r[src].strip(), int(r[ex] or 0), float(r[si] or 0)) for k, r in enumerate(rows[2:]) if len(r) > si]
    tot = sum(d[3] for d in data)
    print(rows[0][1][:90], "| samples", int(tot))
    # regions delimited by TRYWAIT / UTCBAR / BAR / EXIT
    cur, acc, start = "start", 0.0, 0
    for k, s, e, v in data:
        acc += v
        if any(t in s for t in ("TRYWAIT", "UTCBAR", "BAR.SYNC", "EXIT", "ARRIVE")):
            if acc / tot > 0.004:
                print(f"  lines {start:5d}-{k:5d} {100 * acc / tot:5.1f}%  ends at: {s[:70]}  exec {e}")
            acc, start = 0.0, k + 1
    if len(sys.argv) > 2:
        r = list(csv.reader(open(sys.argv[2])))
        want = ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum")
        for i, n in enumerate(r[0]):
            if n in want:
                print(f"  {n:70s} {r[2][i]} {r[1][i]}")


if __name__ == "__main__":
    main()
