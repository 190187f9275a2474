#!/bin/bash
AB_REPS=2 python scripts/dev/ab_build.py > gpurun_out/prof_sort_plain.log 2>&1 &&
AB_REPS=1 ncu --set full --clock-control none --import-source on -k regex:'onesweep_kernel|hist_kernel' -s 5 -c 3 -f -o gpurun_out/sort_r1 python scripts/dev/ab_build.py > gpurun_out/prof_sort_ncu.log 2>&1
echo rc=$?
