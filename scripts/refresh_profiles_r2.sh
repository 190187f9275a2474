#!/bin/bash
# Turn the files a `scripts/r2_profile.sh TAG` run (+ tests, configs) left in gpurun_out/ into the committed
# summaries under profiles/ (run here, no GPU needed).  usage: scripts/refresh_profiles_r2.sh TAG
set -e
TAG=${1:?tag}; R=r2
cd "$(dirname "$0")/.."
python scripts/ncu_summary.py gpurun_out/trace_$TAG.ncu-rep > profiles/${R}_trace_packet_ncu_full_2p23.json
python scripts/ncu_summary.py gpurun_out/trace20_$TAG.ncu-rep > profiles/${R}_trace_packet_ncu_full_2p20.json
ncu -i gpurun_out/trace_$TAG.ncu-rep --page source --print-source cuda,sass --csv > /tmp/trace_src_$TAG.csv 2>/dev/null
{ echo "# ncu --set full source-level hot spots, trace_packet_kernel<cumulative,32> packet launch (work stealing) of trace_cumulative_sph, 2^24 particles x 2^23 rays";
  python scripts/ncu_lines.py /tmp/trace_src_$TAG.csv 40 0; } > profiles/${R}_trace_packet_hotspots.txt
python - "$TAG" <<'PY'
import json, sys, hashlib, os
tag = sys.argv[1]
d = json.load(open("profiles/r2_trace_packet_ncu_full_2p23.json"))
def num(s):
    p = s.split(); v = float(p[0].replace(",", "")); u = p[1] if len(p) > 1 else ""
    return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u, 1)
e = d[0]      # the packet launch: the dominant kernel
h = hashlib.sha1()
for f in ("trace_packet.cuh", "trace.cu"):
    h.update(open(os.path.join("grace-devel_b200", "csrc", f), "rb").read())
r, w = num(e["dram__bytes_read.sum"]), num(e["dram__bytes_write.sum"])
json.dump({"dram_bytes_per_launch": r + w, "dram_bytes_read": r, "dram_bytes_write": w, "rays_per_launch": 1 << 23,
           "kernel_source_sha": os.environ.get("KERNEL_SHA") or h.hexdigest()[:16],
           "from": "profiles/r2_trace_packet_ncu_full_2p23.json (ncu --set full, launch 0 = trace_packet_kernel<cumulative,32> "
                   "with work stealing, 2^24 particles, 2^23 rays)"},
          open("profiles/trace_traffic.json", "w"), indent=1)
PY
[ -f gpurun_out/build_$TAG.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/build_$TAG.ncu-rep > profiles/${R}_build_kernels_ncu_full.json
[ -f gpurun_out/lists_$TAG.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/lists_$TAG.ncu-rep > profiles/${R}_hit_lists_ncu_full.json
[ -f gpurun_out/hit_lists_$TAG.jsonl ] && grep '^{' gpurun_out/hit_lists_$TAG.jsonl > profiles/${R}_hit_lists_one_vs_two_traversals.jsonl
[ -f gpurun_out/build_times_$TAG.json ] && grep '^{' gpurun_out/build_times_$TAG.json > profiles/${R}_build_times.json
cp gpurun_out/launches_$TAG.csv profiles/${R}_launches.csv
cp gpurun_out/bench_$TAG.json profiles/${R}_bench_n1.json
[ -f gpurun_out/bench_ref_$TAG.json ] && cp gpurun_out/bench_ref_$TAG.json profiles/${R}_bench_reference_arm.json
[ -f gpurun_out/configs_a_$TAG.json ] && cp gpurun_out/configs_a_$TAG.json profiles/${R}_configs.json
for f in gpurun_out/reference_gate_*.txt; do [ -f "$f" ] && cp "$f" profiles/${R}_$(basename $f); done
echo "profiles/ refreshed from gpurun_out/*_$TAG.*"
