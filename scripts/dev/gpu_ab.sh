#!/bin/bash
# usage: gpu_ab.sh tag variant...   ("base" = in-tree lib)
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "base" ]; then unset GRACE_B200_LIB; else export GRACE_B200_LIB=$PWD/grace-devel_b200/variants/libgrace_b200_$v.so; fi
  timeout 300 python scripts/dev/ab_trace.py 2>&1 | tail -1 | tee -a gpurun_out/ab_$TAG.log
done
