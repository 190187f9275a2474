#!/bin/bash
# Round 2 profiling pass on one B200 (gpurun): bench line, its ncu launch list, full ncu captures of the
# traversal kernel (packet launch + fold launch) at the bench workload and at 2^20 rays.
TAG=${1:-r2a}
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-config5 > gpurun_out/ncu_l_$TAG.log 2>&1; echo "launch list rc=$?"
python scripts/dev/probe_one.py 23 cum 2 > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:trace_packet_kernel -s 2 -c 2 -f -o gpurun_out/trace_$TAG \
    python scripts/dev/probe_one.py 23 cum 2 > gpurun_out/ncu_f_$TAG.log 2>&1; echo "ncu full 2^23 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_packet_kernel -s 2 -c 2 -f -o gpurun_out/trace20_$TAG \
    python scripts/dev/probe_one.py 20 cum 2 > gpurun_out/ncu_f20_$TAG.log 2>&1; echo "ncu full 2^20 rc=$?"
ls -la gpurun_out/*_$TAG*
