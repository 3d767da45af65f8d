#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time share of ONE training step
(the launches between the last two `q_xt_kernel` launches).  Usage: python tools/ncu_summary.py launches.csv [out.md]"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("ud::", "").replace("void ", "")
    name = re.sub(r"at::native::.*?(\w+)_kernel.*", r"torch:\1", name)
    return name[:110]


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((int(r["ID"]), r["Kernel Name"], v * scale))
    # one step = from a q_xt launch (first kernel of compute_loss) up to the next one; the streamed optimizer's kernels of
    # that step sit between its backward and the next q_xt.  bench.py --steps 1 --warmup 3 launches q_xt 5 times
    # (3 warm-up, 1 device-timed, 1 end-to-end-timed): the device-timed step is the segment between the last two.
    idx = [i for i, (_, n, _) in enumerate(rows) if "q_xt_kernel" in n]
    if len(idx) >= 2:
        seg = rows[idx[-2]: idx[-1]]
    else:
        seg = rows
    agg, cnt = defaultdict(float), defaultdict(int)
    for _, n, us in seg:
        k = short(n)
        agg[k] += us
        cnt[k] += 1
    total = sum(agg.values())
    out = [f"# kernel time share of one training step ({len(seg)} launches, {total/1e3:.2f} ms serialized, cold-cache ncu times)", "",
           "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, us in sorted(agg.items(), key=lambda kv: -kv[1])[:40]:
        out.append(f"| `{k}` | {cnt[k]} | {us/1e3:.3f} | {100*us/total:.1f}% |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
