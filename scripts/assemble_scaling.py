#!/usr/bin/env python
"""profiles/<round>_scaling_1_2_4_8.json from the bench lines of `bench.py --gpus N` runs:
   python scripts/assemble_scaling.py OUT.json N1.json N2.json N4.json N8.json   (any subset, N=1 first)"""
import json, sys
out, files = sys.argv[1], sys.argv[2:]
runs = []
for f in files:
    line = [l for l in open(f) if l.startswith("{")][-1]
    d = json.loads(line)
    r = {"n_gpus": d["n_gpus"], "mrays_s": d["value"], "ms_per_step": d["ms_per_step"], "steps": d["steps"], "warmup": d["warmup"],
         "e2e_mrays_s": d["e2e"]["value"], "e2e_ms_per_step": d["e2e"].get("ms_per_step"),
         "parity": d.get("parity"), "config5": d.get("config5"), "clocks": d.get("clocks"), "scaling": d.get("scaling")}
    runs.append(r)
base = runs[0]
for r in runs:
    r["speedup_vs_1"] = r["mrays_s"] / base["mrays_s"] * base["n_gpus"]
    r["efficiency"] = r["speedup_vs_1"] / r["n_gpus"]
    r["e2e_speedup_vs_1"] = r["e2e_mrays_s"] / base["e2e_mrays_s"] * base["n_gpus"]
    if r.get("config5") and base.get("config5"):
        r["config5_end_to_end_speedup_vs_1"] = base["config5"]["end_to_end_ms"] / r["config5"]["end_to_end_ms"]
json.dump({"what": "bench.py --gpus N (torchrun, one rank per GPU), strong scaling: one fixed set of 2^23 rays on 2^24 particles at "
                   "every N; builder-measured on one multi-GPU B200 box (gpurun)", "runs": runs}, open(out, "w"), indent=1)
for r in runs:
    print(r["n_gpus"], round(r["mrays_s"], 1), "Mrays/s", round(r["efficiency"], 3), "e2e", round(r["e2e_mrays_s"], 1),
          "config5 x%.2f" % r.get("config5_end_to_end_speedup_vs_1", 0))
