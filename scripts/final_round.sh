#!/bin/bash
# The round's closing measurements on one B200 (gpurun): smoke, GPU tests, bench (+ launch list +
# full ncu capture of the trace kernel), per-config table, comparisons with the reference CUDA build.
TAG=${1:-r1f}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/gpu_round.sh $TAG
timeout 900 python scripts/bench_configs.py --configs 1,2,3,4,6 > gpurun_out/configs_a_$TAG.json 2> gpurun_out/configs_a_$TAG.err; echo "configs a rc=$?"
timeout 600 python scripts/bench_configs.py --configs 5 > gpurun_out/configs_b_$TAG.json 2> gpurun_out/configs_b_$TAG.err; echo "configs b rc=$?"
timeout 900 python scripts/compare_reference_cuda.py 24 20 5 > gpurun_out/compare_$TAG.log 2>&1; echo "compare rc=$?"
timeout 900 python scripts/compare_lists.py 24 17 3 > gpurun_out/compare_lists_$TAG.log 2>&1; echo "compare lists rc=$?"
[ -n "${SKIP_NCU:-}" ] && exit 0
AB_REPS=2 python scripts/dev/ab_build.py > gpurun_out/build_plain_$TAG.log 2>&1 &&
AB_REPS=1 ncu --set full --clock-control none --import-source on -k regex:'leaves_kernel|nodes_kernel|onesweep_kernel|morton_keys_kernel|gather16' -s 7 -c 8 -f -o gpurun_out/build_$TAG python scripts/dev/ab_build.py > gpurun_out/build_ncu_$TAG.log 2>&1
echo "build ncu rc=$?"
