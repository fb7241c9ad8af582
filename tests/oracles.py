"""Loaders for the two CPU checkers (test infrastructure; never imported by the product).

  port  oracle/libcxxdes_oracle.so    coroutine-free restatement, builds anywhere
  ref   oracle/_ref/libcxxdes_ref.so  models on the unmodified reference engine
                                      (exists only where /root/reference was present at build time)
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from libcxxdes_b200 import abi

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"


class Oracle:
    def __init__(self, path, prefix):
        self.path = str(path)
        self.lib = C.CDLL(self.path)
        self.prefix = prefix
        self._run = getattr(self.lib, prefix + "_run")
        self._run.restype = C.c_int
        self._run.argtypes = [C.c_int, C.c_void_p, C.POINTER(abi.Clock), C.c_uint64, C.c_uint64,
                              C.c_uint64, C.c_int, C.c_int64, C.POINTER(abi.Columns)]
        self._trace = getattr(self.lib, prefix + "_trace")
        self._trace.restype = C.c_uint64
        self._trace.argtypes = [C.c_int, C.c_void_p, C.POINTER(abi.Clock), C.c_uint64, C.c_uint64,
                                C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.POINTER(C.c_uint64)]
        self._threads = getattr(self.lib, prefix + "_hardware_threads")
        self._threads.restype = C.c_int

    def hardware_threads(self):
        return int(self._threads())

    def run(self, model, params, clock, seed, first_rep, n_reps, threads=1, until=-1):
        cols = abi.HostColumns(n_reps)
        st = cols.as_struct()
        rc = self._run(model, C.addressof(params), C.byref(clock), seed, first_rep, n_reps,
                       threads, until, C.byref(st))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}_run failed with {rc}")
        return cols

    def trace(self, model, params, clock, seed, rep, max_events=4096, until=-1):
        t = np.zeros(max_events, np.int64)
        p = np.zeros(max_events, np.int64)
        k = np.zeros(max_events, np.uint32)
        pr = np.zeros(max_events, np.uint32)
        nd = C.c_uint64()
        total = self._trace(model, C.addressof(params), C.byref(clock), seed, rep, until,
                            max_events, t.ctypes.data, p.ctypes.data, k.ctypes.data,
                            pr.ctypes.data, C.byref(nd))
        n = nd.value
        return int(total), list(zip(t[:n].tolist(), p[:n].tolist(), k[:n].tolist(),
                                    pr[:n].tolist()))


def build_oracles():
    """`make -C oracle` (idempotent): the port always, _ref only where the reference exists."""
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "all"], check=True,
                   stdout=subprocess.DEVNULL)


def port_oracle():
    path = ORACLE_DIR / "libcxxdes_oracle.so"
    if not path.exists():
        build_oracles()
    return Oracle(path, "cxxdes_oracle")


def ref_oracle():
    """The reference-engine oracle, or None when it was never built (no /root/reference)."""
    path = ORACLE_DIR / "_ref" / "libcxxdes_ref.so"
    if not path.exists() and os.path.isdir("/root/reference/include/cxxdes"):
        build_oracles()
    if not path.exists():
        return None
    return Oracle(path, "cxxdes_ref")


MM1_CLOCK = abi.Clock.of(unit=(1, abi.SECONDS), precision=(1, abi.MILLISECONDS))
