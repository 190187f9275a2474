#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` by CUDA source line
(instructions executed, stall samples, lanes per instruction).  usage: ncu_lines.py file.csv [top] [launch]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
# sections: (file, function) -> rows ; launches repeat the same (file, function) sequence
sections = []; fpath = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name": sections.append({"file": fpath, "fn": r[1], "rows": []}); continue
    if sections: sections[-1]["rows"].append(r)
# group launches: a new launch starts when a (file) repeats
launches = [[]]; seen = set()
for s in sections:
    if s["file"] in seen: launches.append([]); seen = set()
    seen.add(s["file"]); launches[-1].append(s)
L = launches[which]
agg = {}; tot_i = tot_s = 0
for s in L:
    hdr = s["rows"][0]
    iInst = hdr.index("Instructions Executed"); iSamp = hdr.index("# Samples"); iThr = hdr.index("Thread Instructions Executed")
    for r in s["rows"][1:]:
        if r[0] == "" or len(r) <= iThr: continue
        try: ins = int(r[iInst]); smp = int(r[iSamp]); thr = int(r[iThr])
        except ValueError: continue
        key = (s["file"].split("/")[-1], int(r[0]), r[1].strip()[:100])
        a = agg.setdefault(key, [0, 0, 0]); a[0] += ins; a[1] += smp; a[2] += thr
        tot_i += ins; tot_s += smp
print("launches in report: %d ; launch %d: total inst %.3e, samples %d" % (len(launches), which, tot_i, tot_s))
order = sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]
for (f, ln, src), (ins, smp, thr) in order:
    print("%-18s %4d inst %5.1f%% samp %5.1f%% lanes %4.1f | %s" % (f[:18], ln, 100.0 * ins / tot_i, 100.0 * smp / max(tot_s, 1), thr / max(ins, 1), src))
if len(sys.argv) > 4:
    # region sums: "name:lo-hi,name:lo-hi" over the file given by argv[5] (default trace_packet.cuh)
    fsel = sys.argv[5] if len(sys.argv) > 5 else "trace_packet.cuh"
    print("regions of", fsel)
    for spec in sys.argv[4].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        ins = sum(v[0] for k, v in agg.items() if k[0] == fsel and lo <= k[1] <= hi)
        smp = sum(v[1] for k, v in agg.items() if k[0] == fsel and lo <= k[1] <= hi)
        thr = sum(v[2] for k, v in agg.items() if k[0] == fsel and lo <= k[1] <= hi)
        print("  %-10s inst %5.1f%% samp %5.1f%% lanes %4.1f" % (name, 100.0 * ins / tot_i, 100.0 * smp / tot_s, thr / max(ins, 1)))
    other = sum(v[1] for k, v in agg.items() if k[0] != fsel)
    print("  other files samp %5.1f%%" % (100.0 * other / tot_s))
