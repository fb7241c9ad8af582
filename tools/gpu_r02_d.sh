#!/bin/bash
# Round 2, GPU call D: straight-line ALOHA kernel -- parity and throughput, cheap ncu sections at a representative size.
set -u
O=gpurun_out/r02d
mkdir -p $O
python -m pytest tests/test_models_gpu.py tests/test_models_fuzz.py tests/test_coop_heap.py -m gpu -q -k "aloha or fuzz or sharding" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_gpu.log | tee -a $O/summary.txt
python tools/profile_model.py aloha 262144 3 2>&1 | tee -a $O/summary.txt
ncu --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section InstructionStats --section SourceCounters --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --import-source on --clock-control none -k regex:aloha_fast_kernel -c 1 -o $O/aloha_flat python tools/profile_model.py aloha 113664 1 > $O/ncu_flat.log 2>&1
grep "aloha" $O/ncu_flat.log | tee -a $O/summary.txt
