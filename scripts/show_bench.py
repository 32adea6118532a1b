#!/usr/bin/env python
"""Pretty-print a bench.py JSON line (the per-kernel tables one per row)."""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.strip().startswith("{")][-1])
r = d.pop("roofline", None) or {}
ak, hk = r.pop("all_kernels_ms_per_step", {}), r.pop("hbm_kernels", {})
print(json.dumps(d, indent=1))
print(json.dumps(r, indent=1))
for k, v in ak.items():
    print(f"{k:36s} {v}")
for k, v in hk.items():
    print(f"{k:36s} {v}")
