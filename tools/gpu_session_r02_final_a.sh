#!/bin/bash
# Round-2 final session, part A: the whole -m gpu suite, then ncu captures of the ALOHA and M/M/c first-pass kernels as
# benchmarked (tools/gpu_ncu_models.sh).  Outputs under gpurun_out/<outdir>/.
set -u
O=gpurun_out/$1; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_full.log 2>&1; echo "full pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu_full.log | tee -a $O/summary.txt
bash tools/gpu_ncu_models.sh $1
