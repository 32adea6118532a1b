#!/usr/bin/env python
"""profiles/conv_traffic.json from an `ncu --set full` capture of the conv kernels (bench.py reads it for roofline.traffic):
mean dram__bytes_read.sum + dram__bytes_write.sum per launch of conv3x3_halo_kernel, tagged with the hash of the kernel
source so that bench.py can tell a stale capture from a current one.
Usage: python scripts/ncu_traffic.py gpurun_out/prof_conv3x3_halo.ncu-rep [label]"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


per = []
for d in data:
    if "conv3x3_halo_kernel" not in d[idx["Kernel Name"]]:
        continue
    rd = to_bytes(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
    wr = to_bytes(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    per.append({"kernel": d[idx["Kernel Name"]].split("(")[0][-60:], "grid": d[idx["Grid Size"]], "dram_read": rd, "dram_write": wr,
                "time_us": float(d[idx["gpu__time_duration.sum"]].replace(",", "")) / (1e3 if units[idx["gpu__time_duration.sum"]].startswith("n") else 1),
                "tensor_pipe_pct": d[idx.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)]})
src = os.path.join(ROOT, "enhanced_unet_b200", "csrc", "conv_halo.cu")
rec = {"source": f"ncu --set full, {os.path.basename(rep)}" + (f" ({sys.argv[2]})" if len(sys.argv) > 2 else ""),
       "conv_halo_cu_sha16": hashlib.sha256(open(src, "rb").read()).hexdigest()[:16],
       "launches": len(per), "dram_bytes_per_launch_mean": sum(p["dram_read"] + p["dram_write"] for p in per) / max(1, len(per)),
       "per_launch": per}
json.dump(rec, open(os.path.join(ROOT, "profiles", "conv_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in rec.items() if k != "per_launch"}))
