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
# build kernels and the one-pass hit-list kernels (ncu --set full), configs, reference arm
timeout 900 ncu --set full --clock-control none -k regex:'leaves_|nodes_kernel|onesweep_kernel|morton_keys_kernel|minmax_kernel|deltas_' -s 11 -c 11 -f -o gpurun_out/build_$TAG \
    python scripts/dev/build_once.py > gpurun_out/ncu_b_$TAG.log 2>&1; echo "ncu build rc=$?"
python scripts/dev/build_once.py > gpurun_out/build_times_$TAG.json 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'rec_|trace_packet_kernel<5' -s 4 -c 4 -f -o gpurun_out/lists_$TAG \
    python scripts/dev/lists_once.py > gpurun_out/ncu_h_$TAG.log 2>&1; echo "ncu lists rc=$?"
{ GRACE_B200_ONE_PASS_LISTS=0 python scripts/dev/ab_lists.py | sed 's/"tag": "/"tag": "two_traversals_/'; python scripts/dev/ab_lists.py | sed 's/"tag": "/"tag": "one_traversal_/'; } > gpurun_out/hit_lists_$TAG.jsonl 2>&1
timeout 1200 python scripts/bench_configs.py > gpurun_out/configs_a_$TAG.json 2> gpurun_out/configs_a_$TAG.err; echo "configs rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "reference arm rc=$?"
ls -la gpurun_out/*_$TAG*
