#!/bin/bash
set -u
O=gpurun_out/mm1_profile; mkdir -p $O
python -m pytest tests/test_mm1_gpu.py tests/test_mm1_fuzz.py tests/test_trace_gpu.py -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; cut -c1-250 $O/bench.json | tee -a $O/summary.txt
ncu --set full --import-source on --clock-control none -k regex:mm1_ahead_kernel -c 1 -o $O/mm1_ahead_v5 python tools/profile_model.py mm1 1048576 1 > $O/ncu.log 2>&1
tail -2 $O/ncu.log | tee -a $O/summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
