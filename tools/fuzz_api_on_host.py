#!/usr/bin/env python
"""Random SEQUENCES of C-ABI calls on one handle -- run_until(t), run_for(dt), run(), reset(), fetch / reduce after every
one -- for every model, against what the reference's environment would show after the same calls (CPU oracle).  On the
CPU emulation of the device (tests/native/cuda_on_host; TEST INFRASTRUCTURE), so it exercises the library's own host
logic (which kernel a piece takes, which saved state is valid, the clock after exhaustion) and the kernels' saved state.

Reference semantics restated (environment.ipp:179-214): run_until(t) dispatches what is due up to t and sets now =
max(now, t); run_for(dt) = run_until(now + dt); run() dispatches everything and leaves now where the last event (or an
earlier run_until) put it; after that nothing is left, run_until / run_for only move the clock.

    python tools/fuzz_api_on_host.py [--first 0] [--count 200] [--sms 1]
"""
import argparse
import os
import sys
import traceback
from multiprocessing import Pool
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tests" / "native" / "cuda_on_host"))

S, MS, US = 0, 1, 2


def draw_model(rng):
    from libcxxdes_b200 import abi
    kind = rng.choice(["mm1", "aloha", "mmc", "archsim", "program"])
    clock = [((1, S), (1, MS)), ((1, S), (1, US)), ((2, S), (5, MS))][int(rng.integers(3))]
    flags = 0
    if kind == "mm1":
        p = abi.MM1Params(float(rng.choice([2.0, 5.0, 9.0, 11.0])), 10.0, int(rng.choice([3, 40, 300])))
        flags = int(rng.choice([0, 0, abi.FLAG_MM1_LOCKSTEP, abi.FLAG_GENERIC_ENGINE]))
        model = abi.MODEL_MM1
    elif kind == "aloha":
        p = abi.AlohaParams(int(rng.choice([2, 16, 64])), int(rng.choice([0, 5])), float(rng.uniform(0.2, 3.0)), float(rng.uniform(0.2, 3.0)),
                            float(rng.choice([1.0, 0.25])), 10.0, int(rng.choice([2, 12])))
        flags = int(rng.choice([0, 0, abi.FLAG_WIDE_TOKENS]))
        model = abi.MODEL_ALOHA
    elif kind == "mmc":
        p = abi.MMCParams(float(rng.choice([3.0, 9.0, 14.0, 40.0])), 1.0, float(rng.choice([0.0, 0.05, 0.5])), int(rng.choice([1, 2, 8])), 0,
                          int(rng.choice([5, 60, 300])))
        flags = int(rng.choice([0, 0, abi.FLAG_WIDE_TOKENS]))
        model = abi.MODEL_MMC
    elif kind == "archsim":
        clock = ((1, S), (1, S))
        p = abi.ArchSimParams(int(rng.integers(0, 3)), int(rng.integers(1, 12)), int(rng.choice([3, 16])), int(rng.choice([8, 32])),
                              int(rng.choice([1, 2, 4])), 0, int(rng.choice([10, 120])))
        flags = int(rng.choice([0, 0, abi.FLAG_WIDE_TOKENS]))
        model = abi.MODEL_ARCHSIM
    else:
        import test_programs_fuzz as T
        clock = T.UNIT
        p = T.random_program(int(rng.integers(0, 4000)))
        model = abi.MODEL_PROGRAM
    return kind, model, p, clock, flags


class Session:
    """One handle and what the reference's environment would show after the calls made on it so far."""

    def __init__(self, rng, port, tag):
        import libcxxdes_b200 as L
        from libcxxdes_b200 import abi
        self.rng, self.port, self.abi = rng, port, abi
        self.kind, self.model, p, self.clock, flags = draw_model(rng)
        self.params = p.to_params() if hasattr(p, "to_params") else p
        self.oclock = abi.Clock.of(*self.clock)
        self.n = int(rng.choice([40, 200, 700]))
        self.seed, self.first = int(rng.integers(1 << 40)), int(rng.integers(1 << 33))
        self.full = port.run(self.model, self.params, self.oclock, self.seed, self.first, self.n, threads=2)
        self.skip = bool(self.full.status.any())      # (the oracle itself bounded this generated program: not a case)
        self.t_end = max(1, int(self.full.end_time.max()))
        self.ens = None
        if not self.skip:
            self.ens = L.Ensemble(p, self.n, seed=self.seed, first_replication=self.first, flags=flags)
            self.ens.time_unit(*self.clock[0]).time_precision(*self.clock[1])
        self.log = [f"[{tag}] {self.kind} flags={flags} n={self.n} clock={self.clock}"]
        self.horizon, self.exhausted = -1, False
        self.floor, self.shift = np.zeros(self.n, np.int64), np.zeros(self.n, np.int64)

    def close(self):
        if self.ens is not None:
            self.ens.close()

    def step(self):
        rng, ens, abi, n, full, port = self.rng, self.ens, self.abi, self.n, self.full, self.port
        model, params, oclock, seed, first, kind, log = self.model, self.params, self.oclock, self.seed, self.first, self.kind, self.log
        op = rng.choice(["run_until", "run_until", "run_for", "run_for", "run", "reset"])
        if op == "run_until":
            t = int(rng.uniform(0, 1.3) * self.t_end)
            log.append(f"run_until({t})")
            ens.run_until(t)
            if self.exhausted:
                self.floor = np.maximum(self.floor + self.shift, t); self.shift[:] = 0      # now = max(now, t) per replication
            else:
                self.horizon = max(self.horizon, t)
        elif op == "run_for":
            dt = int(rng.uniform(0, 0.5) * self.t_end)
            log.append(f"run_for({dt})")
            ens.run_for(dt)
            if self.exhausted:
                self.shift = self.shift + dt
            else:
                self.horizon = (0 if self.horizon < 0 else self.horizon) + dt
        elif op == "run":
            log.append("run()")
            ens.run()
            if not self.exhausted:
                self.exhausted = True
                self.floor = np.maximum(full.end_time, self.horizon if self.horizon >= 0 else 0).astype(np.int64)
                self.shift = np.zeros(n, np.int64)
        else:
            log.append("reset()")
            ens.reset()
            self.horizon, self.exhausted = -1, False
            return
        horizon, exhausted = self.horizon, self.exhausted
        got = ens.fetch(strict=False)
        if exhausted:
            want = full
            want_end = self.floor + self.shift
            unfinished = np.zeros(n, bool)
        else:
            want = port.run(model, params, oclock, seed, first, n, threads=2, until=horizon)
            want_end = want.end_time
            unfinished = want.n_events != full.n_events      # (the oracle does not flag them: fewer events than the whole run)
        bad = (got.status & ~np.uint32(abi.REP_UNFINISHED)) != 0
        if kind == "program":
            # what even the interpreter's last pass bounds may flag a replication: never compared
            assert ((got.status[bad] & ~np.uint32(abi.REP_UNFINISHED | abi.REP_HEAP_OVERFLOW | abi.REP_WAITER_OVERFLOW | abi.REP_QUEUE_OVERFLOW)) == 0).all(), "status"
        else:
            assert not bad.any(), f"status {np.unique(got.status[bad])}"
        ok = ~bad
        assert np.array_equal(((got.status & abi.REP_UNFINISHED) != 0)[ok], unfinished[ok]), "unfinished flags"
        assert np.array_equal(got.n_events[ok], want.n_events[ok]), "n_events"
        assert np.array_equal(got.trace_hash[ok], want.trace_hash[ok]), "trace_hash"
        assert np.array_equal(got.end_time[ok], np.asarray(want_end)[ok]), "end_time"
        for k in range(abi.N_F64):
            assert np.array_equal(got.f64[k].view(np.uint64)[ok], want.f64[k].view(np.uint64)[ok]), f"f64[{k}]"
        for k in range(abi.N_I64):
            assert np.array_equal(got.i64[k][ok], want.i64[k][ok]), f"i64[{k}]"
        acc, _ = ens.reduce()      # the on-device reduction against the fetched columns (flagged rows included, as it counts them)
        assert acc.n_replications == n and acc.n_events == int(got.n_events.sum()), "reduce n_events"
        assert acc.n_errors == int(bad.sum()), "reduce n_errors"
        assert acc.trace_hash_sum == int(got.trace_hash.sum(dtype=np.uint64)), "reduce trace_hash_sum"
        assert acc.end_time_sum == int(got.end_time.sum()), "reduce end_time_sum"
        for k in range(abi.N_I64):
            assert acc.i64_sum[k] == int(got.i64[k].sum()), f"reduce i64_sum[{k}]"
        for k in range(abi.N_F64):
            assert abs(acc.f64_sum[k] - float(got.f64[k].sum())) <= 1e-9 * max(1.0, abs(float(got.f64[k].sum()))), f"reduce f64_sum[{k}]"
        if rng.random() < 0.4:
            # the event trace of one replication up to where the handle stands (it must not disturb the handle)
            r = int(rng.integers(n))
            until = -1 if exhausted else horizon
            total, digest, events = ens.trace(r, max_events=256, until=until)
            want_total, want_events = port.trace(model, params, oclock, seed, first + r, max_events=256, until=until)
            log.append(f"trace({r})")
            if ok[r]:
                assert total == want_total == int(want.n_events[r]), "trace: n_events"
                assert [tuple(e) for e in events] == [tuple(e) for e in want_events], "trace: events"
                assert digest == int(want.trace_hash[r]), "trace: digest"


def one_case(case):
    import oracles
    port = oracles.port_oracle()
    rng = np.random.default_rng(31000 + case)
    # one handle, or two handles of (mostly) different models whose calls interleave: nothing of one may leak into the other
    sessions = [Session(rng, port, k) for k in range(1 if rng.random() < 0.5 else 2)]
    sessions = [s for s in sessions if not s.skip]
    kind = "+".join(s.kind for s in sessions) or "none"
    try:
        for _ in range(int(rng.integers(2, 8)) * len(sessions)):
            sessions[int(rng.integers(len(sessions)))].step()
        return case, kind, None
    except BaseException as e:      # noqa: BLE001
        return case, kind, " ;; ".join(" ; ".join(s.log) for s in sessions) + " -> " + "".join(traceback.format_exception_only(type(e), e)).strip()[:400]
    finally:
        for s in sessions:
            s.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--count", type=int, default=200)
    ap.add_argument("--sms", type=int, default=1)
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    args = ap.parse_args()
    import build as cuda_on_host
    os.environ["CXXDES_B200_LIBRARY"] = str(cuda_on_host.build())
    os.environ["CUDA_ON_HOST_SMS"] = str(args.sms)
    import __graft_entry__ as g
    g.build()
    failed, per_kind = [], {}
    with Pool(args.jobs) as pool:
        for case, kind, err in pool.imap_unordered(one_case, range(args.first, args.first + args.count), chunksize=2):
            per_kind[kind] = per_kind.get(kind, 0) + 1
            if err:
                failed.append(case)
                print(f"FAIL case {case}: {err}", flush=True)
    print(f"{args.count - len(failed)} of {args.count} call sequences equal to the oracle ({per_kind}); failed: {sorted(failed)}")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
