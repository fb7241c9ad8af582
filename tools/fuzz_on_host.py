#!/usr/bin/env python
"""More cases of the randomised GPU parity sweeps than the test-suite holds, on the CPU emulation of the device
(tests/native/cuda_on_host; TEST INFRASTRUCTURE -- the kernels' own sources compiled by g++).

The fuzzers of tests/test_mm1_fuzz.py, tests/test_models_fuzz.py and tests/test_programs_fuzz.py draw a case from
`default_rng(base + case)`; the suite runs case 0..13/39, this tool runs any other range of case numbers, every case
against the CPU oracle, bit for bit, on all cores.

    python tools/fuzz_on_host.py [--first 100] [--count 500] [--only aloha,mmc,archsim,mm1,programs,programs_full]
"""
import argparse
import os
import sys
import traceback
from multiprocessing import Pool
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tests" / "native" / "cuda_on_host"))


class _Env:
    """Stand-in for pytest's monkeypatch (setenv only)."""

    def __init__(self):
        self.saved = {}

    def setenv(self, k, v):
        self.saved.setdefault(k, os.environ.get(k))
        os.environ[k] = v

    def undo(self):
        for k, v in self.saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _run(job):
    what, case = job
    import oracles
    port = oracles.port_oracle()
    env = _Env()
    try:
        if what in ("aloha", "mmc", "archsim"):
            import test_models_fuzz as T
            getattr(T, f"test_{what}_fuzz")(True, case)
        elif what == "mm1":
            import test_mm1_fuzz as T
            T.test_gpu_fuzz(port, case)
        else:
            import test_programs_fuzz as T
            T.test_gpu_random_programs_match_the_cpu_interpreter(port, env, case, what == "programs_full")
        return what, case, None
    except BaseException as e:      # noqa: BLE001  (an assertion, or the oracle refusing a generated program)
        return what, case, "".join(traceback.format_exception_only(type(e), e)).strip()[:600]
    finally:
        env.undo()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=100)
    ap.add_argument("--count", type=int, default=200)
    ap.add_argument("--only", default="aloha,mmc,archsim,mm1,programs,programs_full")
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    args = ap.parse_args()
    import build as cuda_on_host
    os.environ["CXXDES_B200_LIBRARY"] = str(cuda_on_host.build())
    os.environ.setdefault("CUDA_ON_HOST_SMS", "2")
    import __graft_entry__ as g
    g.build()
    jobs = [(w, c) for w in args.only.split(",") for c in range(args.first, args.first + args.count)]
    failed = []
    with Pool(args.jobs) as pool:
        for n, (what, case, err) in enumerate(pool.imap_unordered(_run, jobs, chunksize=4), 1):
            if err:
                failed.append((what, case, err))
                print(f"FAIL {what} case {case}: {err}", flush=True)
            if n % 200 == 0:
                print(f"... {n} / {len(jobs)} cases, {len(failed)} failed", flush=True)
    print(f"{len(jobs) - len(failed)} of {len(jobs)} cases equal to the oracle; {len(failed)} failed")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
