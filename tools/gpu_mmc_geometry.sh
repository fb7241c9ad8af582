#!/bin/bash
# M/M/c on narrow tokens: launch geometries / capacities (tools/build_variants.sh variants), with the number of
# replications each hands to the 64-bit kernel.  usage: gpu_mmc_geometry.sh <outdir> variant...
set -u
O=gpurun_out/$1; mkdir -p $O; shift
python -m pytest tests/test_narrow_tokens_gpu.py tests/test_models_gpu.py tests/test_models_fuzz.py tests/test_trace_gpu.py -m gpu -q -x -k "aloha or mmc or fuzz or token" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
cat > /tmp/geo.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import libcxxdes_b200 as L
ens = L.Ensemble(L.MMCParams(9.0, 1.0, 0.5, 8, 0, 1000), 1 << 20, seed=0x5EED)
ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
for _ in range(3):
    ens.run().sync()
    acc, _ = ens.reduce()
    ms = ens.last_run_ms()
    print(f"{ens.kernel_name()}: {ms:.2f} ms, {acc.n_events / ms / 1e6:.2f} Gev/s, errors {acc.n_errors}, hash {acc.trace_hash_sum:#x}, passes {ens.pass_counts()}", flush=True)
PY
echo "== default" | tee -a $O/summary.txt
python /tmp/geo.py 2>&1 | tail -2 | tee -a $O/summary.txt
for v in "$@"; do
  echo "== $v" | tee -a $O/summary.txt
  CXXDES_B200_LIBRARY=$PWD/libcxxdes_b200/_variants/$v.so python /tmp/geo.py 2>&1 | tail -2 | tee -a $O/summary.txt
done
python tools/profile_model.py aloha 262144 2 2>&1 | tail -1 | tee -a $O/summary.txt
