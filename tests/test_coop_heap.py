"""Sub-warp cooperative event heap (libcxxdes_b200/csrc/device/coop_heap.cuh): its steps against std::priority_queue on the
CPU, and the ALOHA kernel built on it (CXXDES_B200_FLAG_COOP_HEAP) against the reference's golden vectors, the oracle and
the thread-per-replication kernel on the GPU -- same tokens in the same order, bit for bit."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

import golden_util
import libcxxdes_b200 as L
from libcxxdes_b200 import abi
from model_cases import MODELS, N_GOLDEN_CASES, clock_of, clock_pairs

ROOT = Path(__file__).resolve().parent.parent


def test_cooperative_heap_steps_match_std_priority_queue(tmp_path):
    """The functions the CUDA group executes (child-vs-child bits, bit walk, resting level by population count, ancestor
    ballot of a push), run lane by lane on the CPU: 9 M random operations with heavy ties pop the same tokens and leave
    the same ARRAY as libstdc++'s std::priority_queue with the reference's comparator (environment.ipp:247-263)."""
    exe = tmp_path / "coop_heap_host"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", f"-I{ROOT / 'include'}", f"-I{ROOT / 'libcxxdes_b200' / 'csrc'}",
                    str(ROOT / "tests" / "native" / "coop_heap_host.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatching runs" in r.stdout


def _run(params, n_reps, clock, seed, first=0, until=-1, flags=0):
    ens = L.Ensemble(params, n_reps, seed=seed, first_replication=first, flags=flags)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    if until < 0:
        ens.run()
    else:
        ens.run_until(until)
    cols = ens.fetch(strict=False)
    acc, _ = ens.reduce()
    name = ens.kernel_name()
    ens.close()
    return cols, acc, name


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(N_GOLDEN_CASES))
def test_gpu_cooperative_aloha_matches_reference_golden(built, idx):
    model, cls, fname = MODELS["aloha"]
    g = golden_util.load(fname)
    case = g["cases"][idx]
    got, _, name = _run(cls(*case["params"]), case["n_replications"], clock_pairs(g), case["seed"],
                        first=case["first_replication"], until=case["run_until"], flags=abi.FLAG_COOP_HEAP)
    assert name == "aloha_coop_kernel"
    assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
    golden_util.assert_columns_equal(got, golden_util.golden_columns(case), what=f"cooperative aloha golden {idx}")


@pytest.mark.gpu
@pytest.mark.parametrize("fields,n_reps", [
    ((64, 50, 0.1, 3.0, 1.0, 100.0, 0), 2000),      # the offered-load sweep, 64 stations
    ((64, 0, 3.0, 3.0, 1.0, 1000.0, 40), 512),      # heavy load: many key ties in a 64-entry heap
    ((24, 0, 1.5, 1.5, 1.0, 1000.0, 60), 1024),
    ((1, 0, 0.5, 0.5, 1.0, 1000.0, 30), 64),
    ((3, 0, 2.0, 2.0, 0.0, 1000.0, 2), 100)])        # zero frame time: everything ties
def test_gpu_cooperative_aloha_matches_oracle(port, fields, n_reps):
    model, cls, fname = MODELS["aloha"]
    g = golden_util.load(fname)
    p = cls(*fields)
    got, acc, _ = _run(p, n_reps, clock_pairs(g), 0xFEED, first=17, flags=abi.FLAG_COOP_HEAP)
    want = port.run(model, p, clock_of(g), 0xFEED, 17, n_reps, threads=os.cpu_count())
    assert not got.status.any() and acc.n_errors == 0
    golden_util.assert_columns_equal(got, want, what=f"cooperative aloha {fields}")
    assert acc.trace_hash_sum == int(want.trace_hash.sum(dtype=np.uint64))


@pytest.mark.gpu
def test_gpu_cooperative_aloha_equals_thread_per_replication_at_size(built):
    """More replications than the grid holds groups (work stealing by groups), both kernels, same columns."""
    model, cls, fname = MODELS["aloha"]
    g = golden_util.load(fname)
    p = cls(64, 50, 0.1, 3.0, 1.0, 20.0, 0)
    a, acc_a, name_a = _run(p, 60000, clock_pairs(g), 5, flags=abi.FLAG_COOP_HEAP)
    b, acc_b, name_b = _run(p, 60000, clock_pairs(g), 5)
    assert (name_a, name_b) == ("aloha_coop_kernel", "aloha_compact_kernel<false>")
    golden_util.assert_columns_equal(a, b, what="cooperative vs thread-per-replication")
    assert acc_a.trace_hash_sum == acc_b.trace_hash_sum and acc_a.n_errors == 0


@pytest.mark.gpu
def test_cooperative_flag_is_aloha_only(built):
    with pytest.raises(L.CxxdesError):
        L.Ensemble(L.MM1Params(5.0, 10.0, 10), 4, flags=abi.FLAG_COOP_HEAP).run()
