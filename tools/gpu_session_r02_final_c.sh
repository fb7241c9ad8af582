#!/bin/bash
# Round-2 final session, part C (after the memory-hierarchy kernel's narrow layout): whole -m gpu suite, smoke, the two
# memory-hierarchy configs through bench.py.  Outputs under gpurun_out/<outdir>/.
set -u
O=gpurun_out/$1; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu_full.log 2>&1; echo "full pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu_full.log | tee -a $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
for c in ${CONFIGS:-archsim archsim4}; do
  python bench.py --config $c --steps 3 --warmup 3 --cpu-seconds 6 > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?" | tee -a $O/summary.txt
  cut -c1-330 $O/bench_$c.json | tee -a $O/summary.txt
done
