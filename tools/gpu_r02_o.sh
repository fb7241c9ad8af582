#!/bin/bash
set -u
O=gpurun_out/r02o; mkdir -p $O
cat > /tmp/ab3.py <<'PY'
import os, sys, json
sys.path.insert(0, '.')
import libcxxdes_b200 as L
for first in (0, 5 << 20):
    ens = L.Ensemble(L.MM1Params(5.0, 10.0, 10000), 1 << 20, seed=0x5EED, first_replication=first)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    for _ in range(4):
        ens.run().sync()
        pc = ens.pass_counts()
        v = pc["helper_longest_ns"]
        print(json.dumps({"shard": first >> 20, "ms": round(ens.last_run_ms(), 3), "helper": pc["helper"], "resident_ms": (v & ((1 << 62) - 1)) / 1e6 if v >> 62 else None, "raw": v}), flush=True)
    ens.close()
PY
python /tmp/ab3.py 2>&1 | tee -a $O/summary.txt
