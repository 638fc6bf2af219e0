#!/usr/bin/env python
"""Prints the headline fields of a bench.py JSON line: python tools/print_bench.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("headline", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 4), "n_gpus", d["n_gpus"])
if "roofline" in d:
    r = d["roofline"]
    print("roofline frac", round(r["frac"], 4), "effective", round(r.get("frac_effective", 0) or 0, 4), "conv share",
          round(r.get("conv_share_of_step", 0) or 0, 4), "traffic", r.get("traffic"))
if d.get("e2e"):
    print("e2e", round(d["e2e"]["value"], 1), d["e2e"].get("h2d_bytes_per_step"), d["e2e"].get("d2h_bytes_per_step"))
op = d.get("early_exit_operating_point")
if op:
    print("operating point", round(op["value"], 1), "e2e", op.get("e2e", {}).get("value") if op.get("e2e") else None)
for k, v in (d.get("extra_workloads") or {}).items():
    if isinstance(v, dict):
        print(k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items()
                  if kk in ("value", "unit", "ms_per_step", "e2e", "baseline_torch", "torch_value", "speedup_vs_torch")})
print("clocks", d.get("clocks"))
print("gpu_launches", d.get("gpu_launches"), "cpu_baseline", (d.get("cpu_baseline") or {}).get("value"))
