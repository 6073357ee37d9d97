#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, mean, share).

    python tools/ncu_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    cols = rows[hdr]
    ki, vi, mi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Name"), cols.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi or "time_duration" not in r[mi]:          # a multi-metric list: only the duration rows
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        scale = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(r[ui], 1.0)
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")) * scale)
    total = sum(sum(v) for v in agg.values())
    n = sum(len(v) for v in agg.values())
    print(f"source: `{path}` -- {n} launches, {total / 1e3:.1f} us summed (ncu per-launch times are cold-cache and")
    print("serialised: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        ours = "**" if ("tsc::" in k or k.startswith("tc::")) else ""
        print(f"| {ours}`{k[:90]}`{ours} | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / total * 100:.1f} % |")


if __name__ == "__main__":
    main(sys.argv[1])
