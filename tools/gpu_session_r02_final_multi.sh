#!/bin/bash
# Round-2 final 8-GPU lines of the configs whose kernels changed late in the round (ALOHA, M/M/c) and of the headline.
set -u
O=gpurun_out/$1; mkdir -p $O
run() {
  local n=$1 out=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline "$@" > $O/$out 2> $O/${out%.json}.err
  echo "$out rc=$? $(grep '^{' $O/$out | cut -c1-230)" | tee -a $O/summary.txt
}
run 8 bench_aloha_8gpu.json --config aloha
run 8 bench_mmc_8gpu.json --config mmc
run 8 bench_weak_8gpu.json
