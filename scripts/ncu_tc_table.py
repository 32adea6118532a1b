#!/usr/bin/env python
"""Per-launch tensor-pipe table from an `ncu --metrics sm__pipe_tc_cycles_active...,sm__pipe_tensor_cycles_active...,
gpu__time_duration.sum --csv` log of scripts/conv_layers_bench.py: consecutive launches of the same kernel / grid are one row.
tc% = tensor-core pipe busy incl. operand fetch from shared memory, tensor% = math cycles."""
import csv
import io
import re
import sys
from collections import OrderedDict

for path in sys.argv[1:]:
    txt = open(path).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    L = OrderedDict()
    for r in rows:
        d = L.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        d[r["Metric Name"].split(".")[0]] = float(r["Metric Value"].replace(",", ""))

    def short(n):
        m = re.match(r"(?:void )?(?:eunet::)?(\w+)(<[^>]*>)?", n)
        return (m.group(1) + (m.group(2) or ""))[:56]

    prev, run, total = None, [], 0.0

    def emit():
        global total
        if not run:
            return
        us = sum(r["gpu__time_duration"] for r in run) / len(run) / 1e3
        total += us
        print(f"{prev[0]:58s} {prev[1]:12s} n={len(run):2d}  tc {sum(r['sm__pipe_tc_cycles_active'] for r in run) / len(run):5.1f} %  "
              f"tensor {sum(r['sm__pipe_tensor_cycles_active'] for r in run) / len(run):5.1f} %  {us:7.1f} us")

    for it in L.values():
        key = (short(it["name"]), it["grid"])
        if key != prev:
            emit()
            run, prev = [], key
        run.append(it)
    emit()
    print(f"sum of per-row mean durations: {total:.0f} us")
