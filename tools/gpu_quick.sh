#!/bin/bash
# quick GPU check: model parity tests + throughput of one config.  usage: gpu_quick.sh <outdir> <model> <reps> [pytest -k expr]
set -u
O=gpurun_out/$1; mkdir -p $O
python -m pytest tests/test_models_gpu.py tests/test_models_fuzz.py -m gpu -q -x -k "${4:-aloha or fuzz}" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
python tools/profile_model.py $2 $3 3 2>&1 | tee -a $O/summary.txt
