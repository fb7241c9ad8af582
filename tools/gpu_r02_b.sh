#!/bin/bash
# Round 2, GPU call B (1 GPU): full -m gpu suite (no -x), half-lane geometries at strong-scaling shard sizes.
set -u
O=gpurun_out/r02b
mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -8 $O/pytest_gpu.log | tee -a $O/summary.txt
cat > /tmp/ab.py <<'PY'
import os, sys, json
sys.path.insert(0, '.')
import libcxxdes_b200 as L
first, n = int(sys.argv[1]), int(sys.argv[2])
ens = L.Ensemble(L.MM1Params(5.0, 10.0, 10000), n, seed=0x5EED, first_replication=first)
ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
ms = []
for _ in range(6):
    ens.run().sync(); ms.append(ens.last_run_ms())
acc, _ = ens.reduce()
print(json.dumps({"first": first, "n": n, "geometry": os.environ.get("CXXDES_B200_MM1_GEOMETRY"),
                  "kernel": ens.kernel_name(), "best_ms": min(ms[1:]), "events_per_s": acc.n_events / (min(ms[1:]) * 1e-3),
                  "passes": ens.pass_counts(), "errors": int(acc.n_errors), "hash_sum": int(acc.trace_hash_sum)}))
PY
for n in 131072 262144 524288 1048576 66304 65536; do for g in 0 1 2; do CXXDES_B200_MM1_GEOMETRY=$g python /tmp/ab.py 0 $n >> $O/ab_geometry.jsonl 2>&1; done; done
cut -c1-260 $O/ab_geometry.jsonl | tee -a $O/summary.txt
