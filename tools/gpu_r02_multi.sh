#!/bin/bash
# Round 2, multi-GPU session (8-GPU box): weak and strong scaling of the headline config at N = 2, 4, 8 through the
# library's own collective; configs 3-5 at N = 8.  One JSON line per run under gpurun_out/r02m/.
set -u
O=gpurun_out/r02m; mkdir -p $O
run() {  # run <N> <outfile> <bench args...>
  local n=$1 out=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline "$@" > $O/$out 2> $O/${out%.json}.err
  echo "$out rc=$? $(cut -c1-260 $O/$out)" | tee -a $O/summary.txt
}
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_weak_1gpu.json 2> $O/bench_weak_1gpu.err; echo "1gpu $(cut -c1-200 $O/bench_weak_1gpu.json)" | tee -a $O/summary.txt
for n in 8 4 2; do run $n bench_weak_${n}gpu.json; done
for n in 8 4 2; do run $n bench_strong_${n}gpu.json --scaling strong; done
run 8 bench_archsim_8gpu.json --config archsim
run 8 bench_aloha_8gpu.json --config aloha
run 8 bench_mmc_8gpu.json --config mmc
run 8 bench_archsim4_8gpu.json --config archsim4
python tests/test_host_cpp.py > /dev/null 2>&1
python -m pytest tests/test_host_cpp.py -m gpu -q 2>&1 | tail -2 | tee -a $O/summary.txt
