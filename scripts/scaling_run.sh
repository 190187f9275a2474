#!/bin/bash
# bench.py at N GPUs on this box (N > 1: torchrun, one rank per GPU).  usage: scripts/scaling_run.sh N TAG [steps]
N=${1:?gpus}; TAG=${2:?tag}; STEPS=${3:-10}
if [ "$N" = 1 ]; then
  timeout 900 python bench.py --gpus 1 --steps $STEPS --warmup 3 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
fi
echo "N=$N rc=$?"; tail -c 600 gpurun_out/bench_n${N}_$TAG.json
