"""ctypes mirror of include/cxxdes_b200.h (the C ABI of the ensemble engine).

Only declarations live here: structures, enum values and a helper that gives
numpy-backed result columns.  Nothing in this module computes anything.
"""
import ctypes as C

import numpy as np

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_CAPACITY = 0, 1, 2, 3, 4
SECONDS, MILLISECONDS, MICROSECONDS, NANOSECONDS, PICOSECONDS = range(5)
MODEL_MM1, MODEL_ALOHA, MODEL_MMC, MODEL_ARCHSIM, MODEL_PROGRAM = 1, 2, 3, 4, 5
N_F64 = 2
N_I64 = 2
REP_QUEUE_OVERFLOW, REP_HEAP_OVERFLOW, REP_WAITER_OVERFLOW, REP_TIME_RANGE, REP_UNFINISHED = (
    1, 2, 4, 8, 16)
REP_BAD_PROGRAM = 32
INHERIT32 = 2 ** 31 - 1          # CXXDES_B200_INHERIT in the 32-bit b field of an op
INHERIT64 = 2 ** 63 - 1          # priority_consts::inherit (defs.ipp:35)
(OP_RETURN, OP_DELAY, OP_UNTIL, OP_DELAY_EXP, OP_DELAY_REAL, OP_FRESH, OP_AWAIT, OP_ASYNC, OP_ALL_OF, OP_ANY_OF,
 OP_CHILD_PROC, OP_CHILD_DELAY, OP_CHILD_UNTIL, OP_Q_PUT, OP_Q_POP, OP_SEM_DOWN, OP_SEM_UP, OP_MTX_ACQUIRE,
 OP_MTX_RELEASE, OP_EV_WAIT, OP_EV_WAKE, OP_LOOP, OP_END_LOOP, OP_SET_REG_NOW, OP_STAT_SECONDS, OP_STAT_TICKS,
 OP_COUNT, OP_MARK, OP_LAZY_TIMEOUT, OP_UNTIL_REG, OP_IF, OP_JUMP, OP_CHILD_ALL_OF, OP_CHILD_ANY_OF, OP_CHILD_EV_WAIT) = range(35)
(SRC_COUNTER, SRC_QUEUE_SIZE, SRC_SEM_VALUE, SRC_NOW, SRC_REGISTER, SRC_MUTEX) = range(6)
CMP = {"==": 0, "!=": 1, "<": 2, "<=": 3, ">": 4, ">=": 5}
FLAG_GENERIC_ENGINE = 1
FLAG_MM1_LOCKSTEP = 2
FLAG_COOP_HEAP = 4
FLAG_WIDE_TOKENS = 8


class Clock(C.Structure):
    """cxxdes_b200_clock: env.time_unit(x) / env.time_precision(x), environment.ipp:43-70."""
    _fields_ = [("unit_t", C.c_int64), ("unit_u", C.c_int32), ("precision_u", C.c_int32),
                ("precision_t", C.c_int64)]

    @classmethod
    def of(cls, unit=(1, SECONDS), precision=(1, SECONDS)):
        return cls(unit[0], unit[1], precision[1], precision[0])


class MM1Params(C.Structure):
    """producer_consumer_example(lambda, mu, n_packets), examples/producer_consumer.cpp:10-20."""
    _fields_ = [("lambda_", C.c_double), ("mu", C.c_double), ("n_packets", C.c_uint64)]


class AlohaParams(C.Structure):
    """aloha_config + the offered-load sweep of main(), examples/aloha.cpp:13-18,87-103."""
    _fields_ = [("num_stations", C.c_uint32), ("sweep_steps", C.c_uint32),
                ("lambda_begin", C.c_float), ("lambda_end", C.c_float),
                ("frame_time", C.c_float), ("packets_scale", C.c_float),
                ("packets_per_station", C.c_uint64)]


class MMCParams(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("mu", C.c_double), ("patience", C.c_double),
                ("servers", C.c_uint32), ("reserved", C.c_uint32), ("n_customers", C.c_uint64)]


class ArchSimParams(C.Structure):
    _fields_ = [("l2_latency", C.c_int64), ("l1_latency", C.c_int64),
                ("l2_capacity", C.c_uint32), ("addr_space", C.c_uint32),
                ("n_requesters", C.c_uint32), ("reserved", C.c_uint32), ("n_loads", C.c_uint64)]


class Op(C.Structure):
    """cxxdes_b200_op"""
    _fields_ = [("code", C.c_uint16), ("a", C.c_uint16), ("b", C.c_int32), ("imm", C.c_int64)]


class Process(C.Structure):
    """cxxdes_b200_process"""
    _fields_ = [("entry", C.c_uint32), ("ordinal", C.c_uint32), ("priority", C.c_int64), ("latency", C.c_int64),
                ("return_latency", C.c_int64), ("return_priority", C.c_int64)]


class ProgramParams(C.Structure):
    """cxxdes_b200_program (host pointers)."""
    _fields_ = [("ops", C.POINTER(Op)), ("procs", C.POINTER(Process)), ("roots", C.POINTER(C.c_uint32)),
                ("queue_max", C.POINTER(C.c_uint64)), ("sem_init", C.POINTER(C.c_uint64)),
                ("sem_max", C.POINTER(C.c_uint64)), ("rates", C.POINTER(C.c_double)),
                ("n_ops", C.c_uint32), ("n_procs", C.c_uint32), ("n_roots", C.c_uint32), ("n_queues", C.c_uint32),
                ("n_sems", C.c_uint32), ("n_mutexes", C.c_uint32), ("n_events", C.c_uint32), ("n_rates", C.c_uint32),
                ("sweep_steps", C.c_uint32), ("reserved", C.c_uint32)]


class Columns(C.Structure):
    _fields_ = [("end_time", C.c_void_p), ("n_events", C.c_void_p), ("trace_hash", C.c_void_p),
                ("status", C.c_void_p), ("f64", C.c_void_p * N_F64), ("i64", C.c_void_p * N_I64)]


class Accumulators(C.Structure):
    _fields_ = [("n_replications", C.c_uint64), ("n_events", C.c_uint64),
                ("n_errors", C.c_uint64), ("trace_hash_sum", C.c_uint64),
                ("end_time_sum", C.c_int64), ("i64_sum", C.c_int64 * N_I64),
                ("f64_sum", C.c_double * N_F64), ("f64_sumsq", C.c_double * N_F64)]


class Config(C.Structure):
    _fields_ = [("model", C.c_int32), ("device", C.c_int32), ("seed", C.c_uint64),
                ("first_replication", C.c_uint64), ("n_replications", C.c_uint64),
                ("clock", Clock), ("params", C.c_void_p), ("params_size", C.c_size_t),
                ("stream", C.c_void_p), ("flags", C.c_uint32), ("queue_capacity", C.c_uint32)]


PARAMS_OF_MODEL = {MODEL_MM1: MM1Params, MODEL_ALOHA: AlohaParams, MODEL_MMC: MMCParams,
                   MODEL_ARCHSIM: ArchSimParams, MODEL_PROGRAM: ProgramParams}


class HostColumns:
    """Per-replication result columns in host memory (numpy), structure of arrays."""

    def __init__(self, n, pinned=False):
        self.n = int(n)
        self._keep = []

        def alloc(dtype):
            if pinned:
                import torch  # pinned host memory is plumbing torch already provides
                raw = torch.empty(self.n * np.dtype(dtype).itemsize, dtype=torch.uint8).pin_memory()
                self._keep.append(raw)
                return raw.numpy().view(dtype)
            return np.zeros(self.n, dtype=dtype)

        self.end_time = alloc(np.int64)
        self.n_events = alloc(np.uint64)
        self.trace_hash = alloc(np.uint64)
        self.status = alloc(np.uint32)
        self.f64 = [alloc(np.float64) for _ in range(N_F64)]
        self.i64 = [alloc(np.int64) for _ in range(N_I64)]

    def nbytes(self):
        return (self.end_time.nbytes + self.n_events.nbytes + self.trace_hash.nbytes +
                self.status.nbytes + sum(a.nbytes for a in self.f64) +
                sum(a.nbytes for a in self.i64))

    def as_struct(self):
        return Columns(self.end_time.ctypes.data, self.n_events.ctypes.data,
                       self.trace_hash.ctypes.data, self.status.ctypes.data,
                       (C.c_void_p * N_F64)(*[a.ctypes.data for a in self.f64]),
                       (C.c_void_p * N_I64)(*[a.ctypes.data for a in self.i64]))
