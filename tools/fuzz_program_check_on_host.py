#!/usr/bin/env python
"""What cxxdes_b200_check_program accepts must be safe to run: generated process programs with a few operation fields
overwritten at random (operands, targets, counts, process entries) go through the check, and every program it ACCEPTS
is run on the CPU emulation of the device built with AddressSanitizer + UBSan (tests/native/cuda_on_host --asan; TEST
INFRASTRUCTURE) -- both interpreters, a bounded horizon, a time limit.  A run may end with any status (REP_BAD_PROGRAM,
capacity flags, unfinished); it must not touch memory it does not own, and the kernels' own guards must stop it.

Run under the sanitizer's preload (the script re-executes itself with it):
    python tools/fuzz_program_check_on_host.py [--first 0] [--count 300]
"""
import argparse
import os
import random
import subprocess
import sys
from multiprocessing import Pool
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tests" / "native" / "cuda_on_host"))

CHILD = r"""
import os, sys, random
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import ctypes as C
import numpy as np
import libcxxdes_b200 as L
from libcxxdes_b200 import abi
import test_programs_fuzz as T
case = int(sys.argv[1])
rng = random.Random(77000 + case)
prog = T.random_program(rng.randrange(4000))
prog.to_params()
n_ops, n_procs = len(prog._ops), len(prog._procs)
interesting = [0, 1, 2, 3, 7, 8, 15, 16, 23, 24, 63, 64, 71, 72, 79, 80, 255, 256, 65535, 65536, -1, -2, n_ops - 1, n_ops, n_ops + 1, n_procs - 1,
               n_procs, n_procs + 1, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1, 2 ** 32, 2 ** 63 - 1, -2 ** 63, -2 ** 31]
for _ in range(rng.randint(1, 3)):
    what = rng.random()
    if what < 0.8:
        i = rng.randrange(n_ops)
        op = list(prog._ops[i])
        field = rng.randrange(4)
        v = rng.choice(interesting)
        if field == 0: v = rng.randrange(0, 48)                     # operation code
        elif field == 1: v &= 0xFFFF                                 # u16 operand
        elif field == 2: v = max(-2 ** 31, min(2 ** 31 - 1, v))      # i32 operand
        else: v = max(-2 ** 63, min(2 ** 63 - 1, v))                 # i64 immediate
        op[field] = v
        prog._ops[i] = tuple(op)
    elif what < 0.9:
        p = rng.choice(prog._procs)
        p[rng.choice(["entry", "latency", "return_latency"])] = max(0, rng.choice(interesting)) & 0xFFFFFFFF
    else:
        prog._roots.append(rng.choice(interesting) & 0xFFFFFFFF)
pp = prog.to_params()
lib = L.load_library()
if lib.cxxdes_b200_check_program(C.byref(pp)) != abi.OK:
    print("REJECTED", lib.cxxdes_b200_last_error().decode()[:100])
    sys.exit(0)
for no_fast in ("0", "1"):
    os.environ["CXXDES_B200_PRG_NO_FAST"] = no_fast
    ens = L.Ensemble(prog, 48, seed=case)
    ens.time_unit(*T.UNIT[0]).time_precision(*T.UNIT[1])
    ens.run_until(3000)
    cols = ens.fetch(strict=False)
    ens.run_for(500)
    ens.fetch(strict=False)
    ens.close()
print("ACCEPTED status", sorted(set(int(s) for s in cols.status)))
"""


def one(case):
    try:
        r = subprocess.run([sys.executable, "-c", CHILD.format(root=str(ROOT)), str(case)], capture_output=True, text=True, timeout=60)
    except subprocess.TimeoutExpired:
        return case, "TIMEOUT", ""
    if r.returncode != 0:
        return case, "CRASH", (r.stderr[-1500:] or r.stdout[-300:])
    return case, r.stdout.split()[0] if r.stdout else "?", r.stdout.strip()[:160]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--count", type=int, default=300)
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    args = ap.parse_args()
    if "CXXDES_B200_LIBRARY" not in os.environ:
        import __graft_entry__ as g
        g.build()
        import build as cuda_on_host
        lib = cuda_on_host.build(asan=True)
        asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
        env = dict(os.environ, CXXDES_B200_LIBRARY=str(lib), CUDA_ON_HOST_SMS="1", LD_PRELOAD=asan,
                   ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="halt_on_error=1:abort_on_error=1")
        os.execve(sys.executable, [sys.executable, *sys.argv], env)
    tally, bad = {}, []
    with Pool(args.jobs) as pool:
        for case, outcome, detail in pool.imap_unordered(one, range(args.first, args.first + args.count)):
            tally[outcome] = tally.get(outcome, 0) + 1
            if outcome in ("CRASH", "TIMEOUT", "?"):
                bad.append(case)
                print(f"{outcome} case {case}: {detail}", flush=True)
    print(f"{args.count} mutated programs: {tally}; crashed / timed out: {sorted(bad)}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
