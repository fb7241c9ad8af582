#!/bin/bash
set -u
O=gpurun_out/per_shard; mkdir -p $O
cat > /tmp/ab2.py <<'PY'
import os, sys, json
sys.path.insert(0, '.')
import libcxxdes_b200 as L
out = []
for first in (0, 5 << 20, 1 << 20, 3 << 20, 0, 6 << 20, 7 << 20):
    ens = L.Ensemble(L.MM1Params(5.0, 10.0, 10000), 1 << 20, seed=0x5EED, first_replication=first)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    ms = []
    for _ in range(4):
        ens.run().sync(); ms.append(round(ens.last_run_ms(), 3))
    pc = ens.pass_counts()
    ens.close()
    print(json.dumps({"shard": first >> 20, "ms": ms, "second_pass": pc["second_pass"], "helper": pc["helper"], "helper_longest_ms": pc["helper_longest_ns"] / 1e6}), flush=True)
PY
python /tmp/ab2.py 2>&1 | tee -a $O/summary.txt
