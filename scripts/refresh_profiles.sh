#!/bin/bash
# Turn the files a `scripts/final_round.sh TAG` run left in gpurun_out/ into the committed
# summaries under profiles/ (run here, no GPU needed).  usage: scripts/refresh_profiles.sh TAG [ROUND]
set -e
TAG=${1:?tag}; R=${2:-r1}
cd "$(dirname "$0")/.."
if [ -f gpurun_out/trace_$TAG.ncu-rep ]; then
python scripts/ncu_summary.py gpurun_out/trace_$TAG.ncu-rep > profiles/${R}_trace_packet_ncu_full.json
python scripts/ncu_summary.py gpurun_out/build_$TAG.ncu-rep > profiles/${R}_build_kernels_ncu_full.json
ncu -i gpurun_out/trace_$TAG.ncu-rep --page source --print-source cuda,sass --csv > /tmp/trace_src_$TAG.csv 2>/dev/null
{ echo "# ncu --set full source-level hot spots, trace_packet_kernel<cumulative,32>, launch 0 (packets) of trace_cumulative_sph, 2^24 particles x 2^20 rays";
  python scripts/ncu_lines.py /tmp/trace_src_$TAG.csv 40 0; } > profiles/${R}_trace_packet_hotspots.txt
python - "$R" <<'PY'
import json, sys
R = sys.argv[1]
d = json.load(open(f"profiles/{R}_trace_packet_ncu_full.json"))
def num(s):
    p = s.split(); v = float(p[0].replace(",", "")); u = p[1] if len(p) > 1 else ""
    return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u, 1)
r = sum(num(e["dram__bytes_read.sum"]) for e in d); w = sum(num(e["dram__bytes_write.sum"]) for e in d)
json.dump({"dram_bytes_per_launch": r + w, "dram_bytes_read": r, "dram_bytes_write": w,
           "what": f"sum of dram__bytes_read.sum + dram__bytes_write.sum over the {len(d)} launches of trace_packet_kernel<cumulative,32> in one trace_cumulative_sph call (2^24 particles, 2^20 rays), ncu --set full, profiles/{R}_trace_packet_ncu_full.json"},
          open("profiles/trace_traffic.json", "w"), indent=1)
PY
cp gpurun_out/launches_$TAG.csv profiles/${R}_launches.csv
fi      # (a SKIP_NCU=1 round leaves the ncu summaries of the previous round in place)
cp gpurun_out/bench_$TAG.json profiles/${R}_bench_n1.json
cp gpurun_out/compare_reference_cuda_2p24_2p20.json profiles/${R}_compare_reference_cuda_2p24_2p20.json
tail -1 gpurun_out/compare_lists_$TAG.log | python -m json.tool > profiles/${R}_compare_lists_2p24_2p17.json
python - "$TAG" "$R" <<'PY'
import json, sys
tag, R = sys.argv[1:3]
a = json.load(open(f"gpurun_out/configs_a_{tag}.json")); b = json.load(open(f"gpurun_out/configs_b_{tag}.json"))
a.update({k: v for k, v in b.items() if k.startswith("config")})
json.dump(a, open(f"profiles/{R}_configs.json", "w"), indent=1)
PY
echo "profiles/ refreshed from gpurun_out/*_$TAG.*"
