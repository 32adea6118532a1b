#!/usr/bin/env python
"""Per-launch duration, DRAM bytes, issue utilisation and the five largest warp-stall reasons of an .ncu-rep."""
import csv
import io
import subprocess
import sys

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    seen = set()
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0][:60]
        key = (name, d[idx["Grid Size"]])
        if key in seen:
            continue
        seen.add(key)
        g = lambda k: d[idx[k]] if k in idx else "?"
        print(f"--- {name} grid {d[idx['Grid Size']]} block {d[idx['Block Size']]}")
        print(f"    time {g('gpu__time_duration.sum')} {units[idx['gpu__time_duration.sum']]}  dram rd {g('dram__bytes_read.sum')} {units[idx['dram__bytes_read.sum']]} "
              f"wr {g('dram__bytes_write.sum')} {units[idx['dram__bytes_write.sum']]}  dram% {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')}")
        print(f"    issue_active% {g('smsp__issue_active.avg.pct_of_peak_sustained_active')}  warps_active% {g('sm__warps_active.avg.pct_of_peak_sustained_active')} "
              f"regs {g('launch__registers_per_thread')}  l1tex% {g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed')} lts% {g('lts__throughput.avg.pct_of_peak_sustained_elapsed')} "
              f"sm% {g('sm__throughput.avg.pct_of_peak_sustained_elapsed')}")
        st = sorted(((float(d[idx[s]] or 0), s[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for s in stall), reverse=True)[:5]
        print("    stalls/issue: " + ", ".join(f"{n} {v:.2f}" for v, n in st))
