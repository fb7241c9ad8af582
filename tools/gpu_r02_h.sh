#!/bin/bash
set -u
O=gpurun_out/r02h; mkdir -p $O
ncu --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section InstructionStats --section SourceCounters --import-source on --clock-control none -k regex:mmc_fast_kernel -c 1 -o $O/mmc_v3 python tools/profile_model.py mmc 104192 1 > $O/ncu_mmc.log 2>&1
grep "mmc" $O/ncu_mmc.log | tee -a $O/summary.txt
ncu --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section InstructionStats --section SourceCounters --import-source on --clock-control none -k regex:aloha_fast_kernel -c 1 -o $O/aloha_v5 python tools/profile_model.py aloha 28416 1 > $O/ncu_aloha.log 2>&1
grep "aloha" $O/ncu_aloha.log | tee -a $O/summary.txt
