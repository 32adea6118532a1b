#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel share of one training step.
Usage: python scripts/summarize_launches.py gpurun_out/launches_r1.csv > profiles/launches_r1_summary.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")


def short(n):
    n = n.replace("void ", "")
    if "eunet::" in n:
        n = n[n.index("eunet::"):]
        return n.split("(")[0][:90]
    return "torch:" + n.split("<")[0].split("(")[0][:60]


def ms(r):
    v = float(r[vi].replace(",", ""))
    return v / 1e6 if r[ui].startswith("n") else (v / 1e3 if r[ui].startswith("u") else v)


# steps are delimited by the input-packing kernel (first kernel of every forward)
starts = [i for i, r in enumerate(rows) if "pack_input_kernel" in r[ki]]
print(f"# {len(rows)} launches captured, {len(starts)} forward passes found")
if len(starts) >= 3:
    seg = rows[starts[-2]:starts[-1]]     # one full warm step (fwd + loss + bwd + optimiser)
else:
    seg = rows
tot = collections.defaultdict(lambda: [0, 0.0])
for r in seg:
    t = tot[short(r[ki])]
    t[0] += 1
    t[1] += ms(r)
T = sum(v[1] for v in tot.values())
print(f"# one training step: {len(seg)} launches, {T:.3f} ms summed kernel time (ncu: serialised, cold cache)")
print(f"{'ms':>9} {'share':>6} {'n':>5}  kernel")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.3f} {100 * v[1] / T:5.1f}% {v[0]:5d}  {k}")
