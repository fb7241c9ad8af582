#!/bin/bash
set -u
O=gpurun_out/r02l; mkdir -p $O
cat > /tmp/ab.py <<'PY'
import os, sys, json
sys.path.insert(0, '.')
import libcxxdes_b200 as L
first, n = int(sys.argv[1]), int(sys.argv[2])
ens = L.Ensemble(L.MM1Params(5.0, 10.0, 10000), n, seed=0x5EED, first_replication=first)
ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
ms, longest = [], []
for _ in range(4):
    ens.run().sync(); ms.append(round(ens.last_run_ms(), 3)); longest.append(ens.pass_counts()["helper_longest_ns"] / 1e6)
print(json.dumps({"first": first, "n": n, "no_overlap": os.environ.get("CXXDES_B200_MM1_NO_OVERLAP"), "helper_blocks": os.environ.get("CXXDES_B200_MM1_HELPER_BLOCKS"),
                  "ms": ms, "helper_longest_ms": longest, "passes": ens.pass_counts()}))
PY
python /tmp/ab.py $((5<<20)) $((1<<20)) >> $O/ab.jsonl 2>&1
python /tmp/ab.py $((3<<20)) $((1<<20)) >> $O/ab.jsonl 2>&1
CXXDES_B200_MM1_HELPER_BLOCKS=1 python /tmp/ab.py $((5<<20)) $((1<<20)) >> $O/ab.jsonl 2>&1
python /tmp/ab.py 4042079 1 >> $O/ab.jsonl 2>&1
CXXDES_B200_MM1_NO_OVERLAP=1 python /tmp/ab.py 4042079 1 >> $O/ab.jsonl 2>&1
python /tmp/ab.py 4042079 32 >> $O/ab.jsonl 2>&1
cat $O/ab.jsonl | cut -c1-400 | tee -a $O/summary.txt
