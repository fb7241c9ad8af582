#!/bin/bash
# One-GPU session: full -m gpu suite, default bench, A/B of the second-pass helper and of the launch
# geometries at a strong-scaling shard size, the other configs, ncu launch list.  Outputs under gpurun_out/single/.
set -u
O=gpurun_out/single
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.log | tee -a $O/summary.txt
python bench.py --steps 3 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" | tee -a $O/summary.txt
cut -c1-400 $O/bench_default.json | tee -a $O/summary.txt
# second pass next to the first vs after it, on the shard rank 3 of the weak-scaling run owns (has two such replications)
cat > /tmp/ab.py <<'PY'
import os, sys, json
sys.path.insert(0, '.')
import numpy as np
import libcxxdes_b200 as L
first, n = int(sys.argv[1]), int(sys.argv[2])
ens = L.Ensemble(L.MM1Params(5.0, 10.0, 10000), n, seed=0x5EED, first_replication=first)
ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
ms = []
for _ in range(6):
    ens.run().sync(); ms.append(ens.last_run_ms())
acc, _ = ens.reduce()
print(json.dumps({"first": first, "n": n, "geometry": os.environ.get("CXXDES_B200_MM1_GEOMETRY"), "no_overlap": os.environ.get("CXXDES_B200_MM1_NO_OVERLAP"),
                  "kernel": ens.kernel_name(), "ms": [round(x, 3) for x in ms], "best_ms": min(ms[1:]), "events_per_s": acc.n_events / (min(ms[1:]) * 1e-3),
                  "passes": ens.pass_counts(), "errors": int(acc.n_errors), "hash_sum": int(acc.trace_hash_sum)}))
PY
for v in 0 1; do CXXDES_B200_MM1_NO_OVERLAP=$v python /tmp/ab.py $((3<<20)) $((1<<20)) >> $O/ab_overlap.jsonl 2>&1; done
for g in 0 1 2; do CXXDES_B200_MM1_GEOMETRY=$g python /tmp/ab.py 0 131072 >> $O/ab_geometry.jsonl 2>&1; done
for g in 0 1 2; do CXXDES_B200_MM1_GEOMETRY=$g python /tmp/ab.py 0 262144 >> $O/ab_geometry.jsonl 2>&1; done
CXXDES_B200_MM1_GEOMETRY=0 python /tmp/ab.py 0 113664 >> $O/ab_geometry.jsonl 2>&1
cat $O/ab_overlap.jsonl $O/ab_geometry.jsonl | cut -c1-330 | tee -a $O/summary.txt
python bench.py --scaling strong --replications 131072 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_strong_shard.json 2> $O/bench_strong_shard.err
cut -c1-300 $O/bench_strong_shard.json | tee -a $O/summary.txt
for c in aloha mmc archsim archsim4; do
  python bench.py --config $c --steps 2 --warmup 3 --cpu-seconds 6 > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?" | tee -a $O/summary.txt
  cut -c1-330 $O/bench_$c.json | tee -a $O/summary.txt
done
# launch list of the default bench (serialised, cold cache: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
echo "ncu rc=$?" | tee -a $O/summary.txt
grep -c mm1_ahead $O/bench_launches.csv | tee -a $O/summary.txt
