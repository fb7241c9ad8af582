#!/bin/bash
# ALOHA on 4-byte tokens against 8-byte tokens: parity tests, then config 3 on both.  usage: gpu_aloha_ab.sh <outdir>
set -u
O=gpurun_out/$1; mkdir -p $O
python -m pytest tests/test_aloha_compact_gpu.py tests/test_models_gpu.py tests/test_models_fuzz.py tests/test_trace_gpu.py -m gpu -q -x -k "aloha or fuzz or token" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_gpu.log | tee -a $O/summary.txt
python tools/profile_model.py aloha 262144 3 2>&1 | tee -a $O/summary.txt
CXB_FLAGS=8 python tools/profile_model.py aloha 262144 2 2>&1 | tee -a $O/summary.txt
