#!/bin/bash
AB_REPS=2 python scripts/dev/ab_build.py > gpurun_out/prof_build_plain.log 2>&1 &&
AB_REPS=1 ncu --set full --clock-control none --import-source on -k regex:'leaves_kernel|nodes_kernel' -s 2 -c 2 -f -o gpurun_out/build_r1 python scripts/dev/ab_build.py > gpurun_out/prof_build_ncu.log 2>&1
echo rc=$?; tail -1 gpurun_out/prof_build_plain.log
