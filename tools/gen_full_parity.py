#!/usr/bin/env python3
"""Checksum of checksums of BASELINE configs[1] at FULL size, from the CPU oracle.

Runs every one of the 1 048 576 replications x 10 000 packets (lambda = 5, mu = 10, seed 0x5EED,
unit 1 s, precision 1 ms) on oracle/port -- the coroutine-free restatement, pinned to the
unmodified reference engine by tests/test_oracle_mm1.py and the golden vectors -- and writes the
integer accumulators the device reduction must reproduce exactly:

    tests/golden/mm1_config2_full.json
        n_events_sum, end_time_sum, trace_hash_sum (wrapping 64-bit), packets_sum, max_queue_sum,
        per-block sums (64 blocks of 16 384 replications: a mismatch is localised without a re-run),
        the replications whose queue outgrew the 31-packet shared-memory window of mm1_ahead_kernel
        (the ones the second pass re-runs).

bench.py asserts the device accumulators of its own first step against these constants before it
times anything ("parity_gate"); tests/test_mm1_gpu.py runs windows of global ids around the listed
long-queue replications.  ~2.6e10 events: about 3 minutes on 8 host threads.

    python tools/gen_full_parity.py [--replications N] [--block B] [--out PATH]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

LAMBDA, MU, PACKETS, SEED = 5.0, 10.0, 10000, 0x5EED
M64 = (1 << 64) - 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replications", type=int, default=1 << 20)
    ap.add_argument("--first", type=int, default=0, help="global id of the first replication (a weak-scaling shard: rank << 20)")
    ap.add_argument("--block", type=int, default=16384)
    ap.add_argument("--window", type=int, default=31, help="packets the shared-memory ring window holds")
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "mm1_config2_full.json"))
    ap.add_argument("--merge", nargs="+", metavar="JSON",
                    help="instead of running anything: merge per-shard outputs of this tool (the 1 Mi shards the ranks of a "
                         "weak-scaling bench own) into one file of shard totals, --out tests/golden/mm1_config2_shards.json")
    a = ap.parse_args()
    if a.merge:
        shards = []
        for path in a.merge:
            d = json.loads(Path(path).read_text())
            shards.append({k: d[k] for k in ("first_replication", "n_replications", "n_events_sum", "end_time_sum",
                                             "trace_hash_sum", "packets_sum", "max_queue_sum")})
            shards[-1]["long_queue"] = [q for q in d["long_queue"] if q["max_queue"] >= a.window - 1]
        shards.sort(key=lambda x: x["first_replication"])
        out = {"model": "mm1", "lambda": LAMBDA, "mu": MU, "n_packets": PACKETS, "seed": SEED,
               "clock": "unit 1 s, precision 1 ms", "generator": "tools/gen_full_parity.py (--first rank<<20, then --merge)",
               "oracle": "oracle/port", "shards": shards}
        Path(a.out).write_text(json.dumps(out, indent=1) + "\n")
        print(f"wrote {a.out}: {len(shards)} shards")
        return

    import oracles
    from libcxxdes_b200 import abi
    port = oracles.port_oracle()
    threads = os.cpu_count() or 1
    p = abi.MM1Params(LAMBDA, MU, PACKETS)
    tot = {"n_events_sum": 0, "end_time_sum": 0, "trace_hash_sum": 0, "packets_sum": 0, "max_queue_sum": 0}
    blocks, long_queue = [], []
    t0 = time.perf_counter()
    for first in range(a.first, a.first + a.replications, a.block):
        n = min(a.block, a.first + a.replications - first)
        c = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, SEED, first, n, threads=threads)
        assert not c.status.any()
        b = {"first": first, "n": n,
             "n_events_sum": int(c.n_events.sum(dtype=np.uint64)),
             "end_time_sum": int(c.end_time.sum()),
             "trace_hash_sum": int(c.trace_hash.sum(dtype=np.uint64)),
             "packets_sum": int(c.i64[0].sum()), "max_queue_sum": int(c.i64[1].sum())}
        blocks.append(b)
        for k in tot:
            tot[k] += b[k]
        # queue length + the packets generated ahead must fit the window (mm1_ahead.cu plan: window - 3)
        for r in np.nonzero(c.i64[1] >= a.window - 6)[0]:
            long_queue.append({"replication": first + int(r), "max_queue": int(c.i64[1][r])})
        done = first + n - a.first
        if (done // a.block) % 8 == 0 or done == a.replications:
            dt = time.perf_counter() - t0
            print(f"{done}/{a.replications} replications, {tot['n_events_sum']:.4g} events, {dt:.0f} s", flush=True)
    tot["trace_hash_sum"] &= M64
    out = {"model": "mm1", "lambda": LAMBDA, "mu": MU, "n_packets": PACKETS, "seed": SEED,
           "clock": "unit 1 s, precision 1 ms", "n_replications": a.replications, "first_replication": a.first,
           "oracle": "oracle/port (restatement of environment.ipp:94-146 + producer_consumer.cpp:31-58)",
           "generator": "tools/gen_full_parity.py", **tot, "blocks": blocks, "long_queue": long_queue,
           "seconds": round(time.perf_counter() - t0, 1), "host_threads": threads}
    Path(a.out).write_text(json.dumps(out, indent=1) + "\n")
    print(f"wrote {a.out}: {tot}")


if __name__ == "__main__":
    main()
