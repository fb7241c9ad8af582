#!/bin/bash
# Round-2 final session, part B: whole -m gpu suite, smoke, default bench + reference arm, the other configs through
# bench.py, ncu launch list of the default bench.  Outputs under gpurun_out/<outdir>/.
set -u
O=gpurun_out/$1; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu_full.log 2>&1; echo "full pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu_full.log | tee -a $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
python bench.py --steps 5 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?" | tee -a $O/summary.txt
cut -c1-300 $O/bench_1gpu.json | tee -a $O/summary.txt
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_1gpu_reference_arm.json 2> $O/bench_ref.err; echo "reference arm rc=$?" | tee -a $O/summary.txt
cut -c1-300 $O/bench_1gpu_reference_arm.json | tee -a $O/summary.txt
for c in aloha mmc archsim archsim4; do
  python bench.py --config $c --steps 3 --warmup 3 --cpu-seconds 6 > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?" | tee -a $O/summary.txt
  cut -c1-330 $O/bench_$c.json | tee -a $O/summary.txt
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
echo "ncu rc=$?" | tee -a $O/summary.txt
grep -c mm1_ahead $O/bench_launches.csv | tee -a $O/summary.txt
