#!/usr/bin/env python
"""Per-kernel DRAM traffic of the training step from an ncu CSV with dram__bytes_read.sum / dram__bytes_write.sum /
gpu__time_duration.sum (tools/ncu_round_s4b.sh):  python tools/ncu_traffic.py X.csv STEPS > profiles/X.json
ncu flushes the caches before every kernel (--cache-control all): the numbers are COLD-cache traffic, an upper bound of what
the kernel moves inside the step, where its operands are usually L2-resident."""
import csv
import json
import sys


def main(path, steps):
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    idi = hdr.index("ID")
    per = {}
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].strip()
        val = float(r[vi].replace(",", ""))
        unit = r[ui]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3,
                 "msecond": 1e3}.get(unit, 1)
        d = per.setdefault(name, dict(launch_ids=set(), read=0.0, write=0.0, us=0.0))
        d["launch_ids"].add(r[idi])
        if "bytes_read" in r[mi]:
            d["read"] += val * scale
        elif "bytes_write" in r[mi]:
            d["write"] += val * scale
        elif "time_duration" in r[mi]:
            d["us"] += val * scale
    out = {}
    for name, d in sorted(per.items(), key=lambda kv: -(kv[1]["read"] + kv[1]["write"])):
        n = len(d["launch_ids"])
        out[name] = dict(launches_per_step=n / steps, dram_read_bytes_per_step=d["read"] / steps,
                         dram_write_bytes_per_step=d["write"] / steps, dram_bytes_per_launch=(d["read"] + d["write"]) / n,
                         us_per_step_cold=d["us"] / steps)
    print(json.dumps(dict(source=path, steps=steps, note="cold-cache (ncu --cache-control all) DRAM traffic per kernel name",
                          kernels=out), indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
