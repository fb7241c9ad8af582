#!/bin/bash
set -u
O=gpurun_out/r02p; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" | tee -a $O/summary.txt
cut -c1-300 $O/bench_default.json | tee -a $O/summary.txt
for c in aloha mmc archsim archsim4; do
  python bench.py --config $c --steps 2 --warmup 3 --cpu-seconds 6 > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?" | tee -a $O/summary.txt
  cut -c1-200 $O/bench_$c.json | tee -a $O/summary.txt
done
python __graft_entry__.py smoke 2>&1 | tail -1 | tee -a $O/summary.txt
