#!/bin/bash
set -u
O=gpurun_out/r02j; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -6 $O/pytest_gpu.log | tee -a $O/summary.txt
python tools/profile_model.py mm1prog 262144 2 2>&1 | tee -a $O/summary.txt
