"""Host-side handle on an ensemble of independent libcxxdes simulations running on one B200.

This mirrors the vocabulary of the reference's `simulation<Derived>` / `environment`
(include/cxxdes/core/simulation.hpp:31-85, core/impl/environment.ipp:12-279) for the batched
run path: `time_unit` / `time_precision` before first use, `run()`, `run_until()`, `run_for()`,
`now()`, `now_seconds()` -- every call forwards to the C ABI in include/cxxdes_b200.h.
There is no CPU implementation behind this class: without the CUDA library (or without a
B200) construction raises.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import abi

# CXXDES_B200_LIBRARY points at another build of the same CUDA library (kernel A/B measurements).
_LIB_PATH = Path(os.environ.get("CXXDES_B200_LIBRARY") or Path(__file__).resolve().parent / "_lib" / "libcxxdes_b200.so")
_lib = None


class CxxdesError(RuntimeError):
    """The reference reports misuse by throwing std::runtime_error; this is its counterpart."""

    def __init__(self, code, message):
        super().__init__(f"cxxdes_b200 error {code}: {message}")
        self.code = code


def library_path():
    return str(_LIB_PATH)


def load_library():
    """dlopen the CUDA library (built in-tree by __graft_entry__.build / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "There is no CPU fallback for the ensemble engine.")
    lib = C.CDLL(str(_LIB_PATH))
    H = C.c_void_p
    lib.cxxdes_b200_last_error.restype = C.c_char_p
    lib.cxxdes_b200_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.cxxdes_b200_version.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.cxxdes_b200_create.argtypes = [C.POINTER(abi.Config), C.POINTER(H)]
    lib.cxxdes_b200_check_program.argtypes = [C.POINTER(abi.ProgramParams)]
    lib.cxxdes_b200_run.argtypes = [H]
    lib.cxxdes_b200_run_until.argtypes = [H, C.c_int64]
    lib.cxxdes_b200_run_for.argtypes = [H, C.c_int64]
    lib.cxxdes_b200_sync.argtypes = [H]
    lib.cxxdes_b200_fetch.argtypes = [H, C.POINTER(abi.Columns)]
    lib.cxxdes_b200_device_columns.argtypes = [H, C.POINTER(abi.Columns)]
    lib.cxxdes_b200_reduce.argtypes = [H, C.POINTER(abi.Accumulators), C.POINTER(C.c_void_p)]
    lib.cxxdes_b200_reduce_allgpus.argtypes = [H, C.c_void_p, C.POINTER(abi.Accumulators)]
    lib.cxxdes_b200_last_run_ms.argtypes = [H, C.POINTER(C.c_float)]
    lib.cxxdes_b200_launch_count.argtypes = [H, C.POINTER(C.c_uint64)]
    lib.cxxdes_b200_pass_counts.argtypes = [H, C.POINTER(C.c_uint64 * 4)]
    lib.cxxdes_b200_kernel_name.argtypes = [H, C.c_char_p, C.c_size_t]
    lib.cxxdes_b200_plan_kernel.argtypes = [C.POINTER(abi.Config), C.c_char_p, C.c_size_t]
    lib.cxxdes_b200_trace.argtypes = [H, C.c_uint64, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.cxxdes_b200_reset.argtypes = [H]
    lib.cxxdes_b200_destroy.argtypes = [H]
    for name in ("device_count", "version", "create", "check_program", "run", "run_until", "run_for", "sync", "fetch",
                 "device_columns", "reduce", "reduce_allgpus", "last_run_ms", "launch_count", "pass_counts", "kernel_name", "plan_kernel", "trace", "reset", "destroy"):
        getattr(lib, "cxxdes_b200_" + name).restype = C.c_int
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "cxxdes_b200_device_count", "cxxdes_b200_create", "cxxdes_b200_check_program", "cxxdes_b200_run", "cxxdes_b200_run_until",
    "cxxdes_b200_run_for", "cxxdes_b200_sync", "cxxdes_b200_fetch", "cxxdes_b200_device_columns",
    "cxxdes_b200_reduce", "cxxdes_b200_reduce_allgpus", "cxxdes_b200_last_run_ms", "cxxdes_b200_launch_count", "cxxdes_b200_pass_counts",
    "cxxdes_b200_kernel_name", "cxxdes_b200_plan_kernel", "cxxdes_b200_trace", "cxxdes_b200_reset",
    "cxxdes_b200_destroy", "cxxdes_b200_last_error", "cxxdes_b200_version",
]


def _check(lib, rc):
    if rc != abi.OK:
        raise CxxdesError(rc, lib.cxxdes_b200_last_error().decode())


def device_count():
    lib = load_library()
    n = C.c_int(0)
    _check(lib, lib.cxxdes_b200_device_count(C.byref(n)))
    return n.value


_MODEL_OF_PARAMS = {abi.MM1Params: abi.MODEL_MM1, abi.AlohaParams: abi.MODEL_ALOHA,
                    abi.MMCParams: abi.MODEL_MMC, abi.ArchSimParams: abi.MODEL_ARCHSIM,
                    abi.ProgramParams: abi.MODEL_PROGRAM}


def plan_kernel(params, unit=(1, abi.SECONDS), precision=(1, abi.SECONDS), flags=0):
    """Name of the first-pass kernel an ensemble of this model would run (ALOHA, M/M/c, memory hierarchy: decided from
    the parameters, the clock and the flags alone -- no handle, no device; cxxdes_b200_plan_kernel)."""
    lib = load_library()
    cfg = abi.Config()
    cfg.model = _MODEL_OF_PARAMS[type(params)]
    cfg.n_replications = 1
    cfg.clock = abi.Clock.of(unit, precision)
    cfg.params = C.addressof(params)
    cfg.params_size = C.sizeof(params)
    cfg.flags = int(flags)
    buf = C.create_string_buffer(128)
    _check(lib, lib.cxxdes_b200_plan_kernel(C.byref(cfg), buf, len(buf)))
    return buf.value.decode()


class Ensemble:
    """`n_replications` independent copies of one model; replication r draws from the Philox
    stream keyed by (seed, first_replication + r), so a shard's results do not depend on how
    the ensemble is split across GPUs."""

    def __init__(self, params, n_replications, seed=0x5EED, first_replication=0, device=0,
                 stream=None, flags=0, queue_capacity=0):
        self._lib = load_library()
        self._handle = None
        if hasattr(params, "to_params"):        # a libcxxdes_b200.program.Program
            self._program = params              # keeps the arrays behind the pointers alive
            params = params.to_params()
        self.params = params
        self.model = _MODEL_OF_PARAMS[type(params)]
        self.n_replications = int(n_replications)
        self.seed = int(seed)
        self.first_replication = int(first_replication)
        self.device = int(device)
        self.stream = stream
        self.flags = int(flags)
        self.queue_capacity = int(queue_capacity)
        # environment(): unit = precision = one_second (environment.ipp:19-21)
        self._unit = (1, abi.SECONDS)
        self._precision = (1, abi.SECONDS)
        self._horizon = None

    # -- env.time_unit / env.time_precision: only before first use (environment.ipp:43-65) --
    def time_unit(self, t, u=abi.SECONDS):
        if self._handle is not None:
            raise CxxdesError(abi.ERR_STATE, "cannot call time_unit(time x) on a used environment!")
        self._unit = (int(t), int(u))
        return self

    def time_precision(self, t, u=abi.SECONDS):
        if self._handle is not None:
            raise CxxdesError(abi.ERR_STATE, "cannot call time_precision(time x) on a used environment!")
        self._precision = (int(t), int(u))
        return self

    def _ensure(self):
        if self._handle is not None:
            return
        cfg = abi.Config()
        cfg.model = self.model
        cfg.device = self.device
        cfg.seed = self.seed
        cfg.first_replication = self.first_replication
        cfg.n_replications = self.n_replications
        cfg.clock = abi.Clock.of(self._unit, self._precision)
        cfg.params = C.addressof(self.params)
        cfg.params_size = C.sizeof(self.params)
        cfg.stream = self.stream
        cfg.flags = self.flags
        cfg.queue_capacity = self.queue_capacity
        h = C.c_void_p()
        _check(self._lib, self._lib.cxxdes_b200_create(C.byref(cfg), C.byref(h)))
        self._handle = h

    # -- simulation::run / run_for, environment::run_until (simulation.hpp:51-62, environment.ipp:179-214) --
    def run(self):
        self._ensure()
        _check(self._lib, self._lib.cxxdes_b200_run(self._handle))
        return self

    def run_until(self, t):
        self._ensure()
        _check(self._lib, self._lib.cxxdes_b200_run_until(self._handle, int(t)))
        return self

    def run_for(self, dt):
        self._ensure()
        _check(self._lib, self._lib.cxxdes_b200_run_for(self._handle, int(dt)))
        return self

    def sync(self):
        self._ensure()
        _check(self._lib, self._lib.cxxdes_b200_sync(self._handle))
        return self

    def fetch(self, columns=None, strict=True):
        """Per-replication results into host memory (numpy SoA)."""
        self._ensure()
        cols = columns if columns is not None else abi.HostColumns(self.n_replications)
        st = cols.as_struct()
        rc = self._lib.cxxdes_b200_fetch(self._handle, C.byref(st))
        if rc != abi.OK and (strict or rc != abi.ERR_CAPACITY):
            _check(self._lib, rc)
        return cols

    def now(self):
        """sim.now() of every replication, in ticks (environment.ipp:24-26)."""
        return self.fetch(strict=False).end_time

    def now_seconds(self):
        """sim.now_seconds() of every replication (environment.ipp:34-36, misc/time.hpp:120-125)."""
        conv = [1, 1e-3, 1e-6, 1e-9, 1e-12]
        return (self.now() * self._precision[0]).astype(np.float64) * conv[self._precision[1]]

    def reduce(self):
        """Shard accumulators (host copy) and the device address of the same struct."""
        self._ensure()
        acc = abi.Accumulators()
        dev = C.c_void_p()
        _check(self._lib, self._lib.cxxdes_b200_reduce(self._handle, C.byref(acc), C.byref(dev)))
        return acc, dev.value

    def reduce_async(self):
        """Launch the reduction and return only the device address (no host sync)."""
        self._ensure()
        dev = C.c_void_p()
        _check(self._lib, self._lib.cxxdes_b200_reduce(self._handle, None, C.byref(dev)))
        return dev.value

    def device_columns(self):
        self._ensure()
        st = abi.Columns()
        _check(self._lib, self._lib.cxxdes_b200_device_columns(self._handle, C.byref(st)))
        return st

    def last_run_ms(self):
        ms = C.c_float()
        _check(self._lib, self._lib.cxxdes_b200_last_run_ms(self._handle, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_uint64()
        _check(self._lib, self._lib.cxxdes_b200_launch_count(self._handle, C.byref(n)))
        return n.value

    def trace(self, replication, max_events=4096, until=-1):
        """(n_events, digest, [(time, priority, kind, process ordinal), ...]) of one replication of this shard: the tokens
        environment::step popped, in order -- same shape as the oracle's trace (tests/oracles.py)."""
        self._ensure()
        t = np.zeros(max_events, np.int64)
        p = np.zeros(max_events, np.int64)
        k = np.zeros(max_events, np.uint32)
        pr = np.zeros(max_events, np.uint32)
        nw, total, digest = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self._lib, self._lib.cxxdes_b200_trace(self._handle, int(replication), int(until), int(max_events),
                                                      t.ctypes.data, p.ctypes.data, k.ctypes.data, pr.ctypes.data,
                                                      C.byref(nw), C.byref(total), C.byref(digest)))
        n = nw.value
        return int(total.value), int(digest.value), list(zip(t[:n].tolist(), p[:n].tolist(), k[:n].tolist(), pr[:n].tolist()))

    def kernel_name(self):
        """The first-pass kernel instantiation a plain run() of this handle launches."""
        self._ensure()
        buf = C.create_string_buffer(160)
        _check(self._lib, self._lib.cxxdes_b200_kernel_name(self._handle, buf, len(buf)))
        return buf.value.decode()

    def reduce_allgpus(self, comm, want_host=True):
        """reduce(), then one NCCL all-reduce of the 88-byte accumulator struct over `comm` (an ncclComm_t with one
        rank per shard, e.g. from libcxxdes_b200.nccl) on the handle's stream, in place on the device.  Returns the
        ensemble totals, or None with want_host=False (no host sync: the call only enqueues)."""
        self._ensure()
        acc = abi.Accumulators() if want_host else None
        _check(self._lib, self._lib.cxxdes_b200_reduce_allgpus(self._handle, C.c_void_p(int(comm)),
                                                               C.byref(acc) if want_host else None))
        return acc

    def pass_counts(self):
        """{second_pass, helper, fallback}: replications of the last launch that were redone outside the first-pass
        kernel's compact representation (diagnostics; results do not depend on it)."""
        c = (C.c_uint64 * 4)()
        _check(self._lib, self._lib.cxxdes_b200_pass_counts(self._handle, C.byref(c)))
        return {"second_pass": int(c[0]), "helper": int(c[1]), "fallback": int(c[2]), "helper_longest_ns": int(c[3])}

    def reset(self):
        if self._handle is not None:
            _check(self._lib, self._lib.cxxdes_b200_reset(self._handle))
        return self

    def close(self):
        if self._handle is not None:
            self._lib.cxxdes_b200_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
