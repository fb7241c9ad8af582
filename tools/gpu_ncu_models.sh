#!/bin/bash
# ncu captures (scheduler / warp-state / instruction / source sections, one launch) of the ALOHA and M/M/c first-pass
# kernels at a size that fills the resident lanes.  usage: gpu_ncu_models.sh <outdir> [aloha_reps] [mmc_reps]
set -u
O=gpurun_out/$1; mkdir -p $O
SECTIONS="--section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section InstructionStats --section ComputeWorkloadAnalysis --section SourceCounters --import-source on --clock-control none"
python tools/profile_model.py aloha ${2:-75776} 1 > $O/aloha_plain.log 2>&1; tail -1 $O/aloha_plain.log | tee -a $O/summary.txt
ncu $SECTIONS -k regex:aloha_compact_kernel -c 1 -o $O/aloha_compact python tools/profile_model.py aloha ${2:-75776} 1 > $O/ncu_aloha.log 2>&1; echo "ncu aloha rc=$?" | tee -a $O/summary.txt
tail -2 $O/ncu_aloha.log | tee -a $O/summary.txt
python tools/profile_model.py mmc ${3:-170496} 1 > $O/mmc_plain.log 2>&1; tail -1 $O/mmc_plain.log | tee -a $O/summary.txt
ncu $SECTIONS -k regex:mmc_fast_kernel -c 1 -o $O/mmc_narrow python tools/profile_model.py mmc ${3:-170496} 1 > $O/ncu_mmc.log 2>&1; echo "ncu mmc rc=$?" | tee -a $O/summary.txt
tail -2 $O/ncu_mmc.log | tee -a $O/summary.txt
