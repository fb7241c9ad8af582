"""Helpers to read tests/golden/*.json (vectors produced by the reference engine itself)."""
import json
from pathlib import Path

import numpy as np

from libcxxdes_b200 import abi

GOLDEN = Path(__file__).resolve().parent / "golden"


def load(name):
    return json.loads((GOLDEN / name).read_text())


def golden_columns(case):
    c = case["columns"]
    n = case["n_replications"]
    cols = abi.HostColumns(n)
    cols.end_time[:] = np.array(c["end_time"], np.int64)
    cols.n_events[:] = np.array(c["n_events"], np.uint64)
    cols.trace_hash[:] = np.array([int(x, 16) for x in c["trace_hash"]], np.uint64)
    for k in range(abi.N_F64):
        cols.f64[k][:] = np.array([int(x, 16) for x in c["f64_bits"][k]], np.uint64).view(np.float64)
    for k in range(abi.N_I64):
        cols.i64[k][:] = np.array(c["i64"][k], np.int64)
    return cols


def assert_columns_equal(got, want, what="", check_status=True):
    """Bit-exact comparison of every column (doubles compared as bit patterns)."""
    assert np.array_equal(got.n_events, want.n_events), f"{what}: n_events differ"
    assert np.array_equal(got.end_time, want.end_time), f"{what}: end_time differs"
    assert np.array_equal(got.trace_hash, want.trace_hash), f"{what}: trace_hash differs"
    for k in range(abi.N_F64):
        assert np.array_equal(got.f64[k].view(np.uint64), want.f64[k].view(np.uint64)), f"{what}: f64[{k}] bits differ"
    for k in range(abi.N_I64):
        assert np.array_equal(got.i64[k], want.i64[k]), f"{what}: i64[{k}] differs"
