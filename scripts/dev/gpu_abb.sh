#!/bin/bash
TAG=$1; shift
for v in "$@"; do
  if [ "$v" = "base" ]; then unset GRACE_B200_LIB; else export GRACE_B200_LIB=$PWD/grace-devel_b200/variants/libgrace_b200_$v.so; fi
  timeout 300 python scripts/dev/ab_build.py 2>&1 | tail -1 | tee -a gpurun_out/abb_$TAG.log
done
