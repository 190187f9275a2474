#!/bin/bash
# One gpurun call: GPU parity tests, bench, ncu launch list, full ncu capture of the trace kernel.
set -u
mkdir -p gpurun_out
TAG=${1:-r1b}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
[ -n "${SKIP_NCU:-}" ] && exit 0     # measurements only: the profiled kernels did not change
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-build-timing --no-reference-cuda"
$CMD2 > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trace_packet_kernel -s 12 -c 4 -f -o gpurun_out/trace_$TAG $CMD2 > gpurun_out/ncu_f_$TAG.log 2>&1
echo "ncu full rc=$?"
