#!/bin/bash
# Round 2, GPU call C: the cooperative-heap ALOHA kernel -- parity, throughput against the thread-per-replication
# kernel at config 3's full size, ncu --set full of both at reduced size.
set -u
O=gpurun_out/r02c
mkdir -p $O
python -m pytest tests/test_coop_heap.py tests/test_models_gpu.py -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_gpu.log | tee -a $O/summary.txt
python tools/profile_model.py aloha 262144 3 2>&1 | tee -a $O/summary.txt
CXB_FLAGS=4 python tools/profile_model.py aloha 262144 3 2>&1 | tee -a $O/summary.txt
ncu --set full --import-source on --clock-control none -k regex:aloha_fast_kernel -c 1 -o $O/aloha_fast python tools/profile_model.py aloha 32768 1 > $O/ncu_fast.log 2>&1
CXB_FLAGS=4 ncu --set full --import-source on --clock-control none -k regex:aloha_coop_kernel -c 1 -o $O/aloha_coop python tools/profile_model.py aloha 32768 1 > $O/ncu_coop.log 2>&1
grep "aloha" $O/ncu_fast.log $O/ncu_coop.log | tee -a $O/summary.txt
ls -la $O | tee -a $O/summary.txt
