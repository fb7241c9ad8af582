#!/bin/bash
set -u
O=gpurun_out/r02k; mkdir -p $O
python -m pytest tests/test_mm1_gpu.py tests/test_mm1_fuzz.py -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
cat > /tmp/ab.py <<'PY'
import os, sys, json
sys.path.insert(0, '.')
import libcxxdes_b200 as L
first, n = int(sys.argv[1]), int(sys.argv[2])
ens = L.Ensemble(L.MM1Params(5.0, 10.0, 10000), n, seed=0x5EED, first_replication=first)
ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
ms = []
for _ in range(5):
    ens.run().sync(); ms.append(ens.last_run_ms())
acc, _ = ens.reduce()
print(json.dumps({"first": first, "n": n, "no_overlap": os.environ.get("CXXDES_B200_MM1_NO_OVERLAP"), "kernel": ens.kernel_name(),
                  "ms": [round(x, 3) for x in ms], "passes": ens.pass_counts(), "errors": int(acc.n_errors)}))
PY
for sh in 0 3 5 6; do python /tmp/ab.py $((sh<<20)) $((1<<20)) >> $O/ab_overlap.jsonl 2>&1; done
CXXDES_B200_MM1_NO_OVERLAP=1 python /tmp/ab.py $((3<<20)) $((1<<20)) >> $O/ab_overlap.jsonl 2>&1
cat $O/ab_overlap.jsonl | cut -c1-300 | tee -a $O/summary.txt
