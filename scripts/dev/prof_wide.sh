#!/bin/bash
export AB_MODE=${AB_MODE:-packet_wide}
python scripts/dev/ab_trace.py > gpurun_out/prof_wide_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trace_packet_kernel -s 4 -c 1 -f -o gpurun_out/trace_wide python scripts/dev/ab_trace.py > gpurun_out/prof_wide_ncu.log 2>&1
echo rc=$?; tail -2 gpurun_out/prof_wide_plain.log
