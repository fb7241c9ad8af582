#!/bin/bash
set -u
O=gpurun_out/multi; mkdir -p $O
run() {
  local n=$1 out=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline "$@" > $O/$out 2> $O/${out%.json}.err
  echo "$out rc=$? $(grep '^{' $O/$out | cut -c1-230)" | tee -a $O/summary.txt
}
python -m pytest tests/test_mm1_gpu.py -m gpu -q 2>&1 | tail -2 | tee -a $O/summary.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_weak_1gpu.json 2> $O/bench_weak_1gpu.err; echo "1gpu $(cut -c1-200 $O/bench_weak_1gpu.json)" | tee -a $O/summary.txt
for n in 8 4 2; do run $n bench_weak_${n}gpu.json; done
run 8 bench_strong_8gpu.json --scaling strong
run 8 bench_aloha_8gpu.json --config aloha
run 8 bench_mmc_8gpu.json --config mmc
