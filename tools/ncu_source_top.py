#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv` export (SASS view): python tools/ncu_source_top.py file.csv [n]"""
import csv
import sys

csv.field_size_limit(10 ** 9)


def main():
    path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(open(path)))
    h = rows[1]
    si, src, ex = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")
    data = []
    for k, r in enumerate(rows[2:]):
        try:
            data.append((float(r[si]), k, r[src].strip(), r[ex]))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1.0
    print(rows[0][1][:100], "| total samples", int(tot), "| SASS lines", len(data))
    for v, k, s, e in sorted(data, reverse=True)[:n]:
        print(f"{100*v/tot:5.1f}%  line {k:5d}  exec {e:>9}  {s[:100]}")


if __name__ == "__main__":
    main()
