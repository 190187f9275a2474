#!/bin/bash
AB_REPS=2 python scripts/dev/ab_build.py > gpurun_out/lb_plain.log 2>&1 &&
AB_REPS=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/launches_build.csv python scripts/dev/ab_build.py > gpurun_out/lb_ncu.log 2>&1
echo rc=$?
