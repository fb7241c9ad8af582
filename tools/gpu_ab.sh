#!/bin/bash
# A/B of library variants (tools/build_variants.sh) on the model configs, after the parity tests of the default build.
# usage: gpu_ab.sh <outdir> "<pytest -k expr>" "<model reps> ..." [variant ...]       e.g. gpu_ab.sh r02x "aloha or mmc" "aloha 262144" "mmc 1048576" loop
set -u
O=gpurun_out/$1; mkdir -p $O; K="$2"; shift 2
python -m pytest tests/test_narrow_tokens_gpu.py tests/test_models_gpu.py tests/test_models_fuzz.py tests/test_trace_gpu.py -m gpu -q -x -k "$K" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -4 $O/pytest_gpu.log | tee -a $O/summary.txt
CONFIGS=(); VARIANTS=()
for a in "$@"; do case "$a" in *" "*) CONFIGS+=("$a");; *) VARIANTS+=("$a");; esac; done
for c in "${CONFIGS[@]}"; do
  echo "== default: $c" | tee -a $O/summary.txt
  python tools/profile_model.py $c 3 2>&1 | tail -2 | tee -a $O/summary.txt
  for v in "${VARIANTS[@]}"; do
    echo "== variant $v: $c" | tee -a $O/summary.txt
    CXXDES_B200_LIBRARY=$PWD/libcxxdes_b200/_variants/$v.so python tools/profile_model.py $c 3 2>&1 | tail -2 | tee -a $O/summary.txt
  done
done
