"""Real C++20 coroutine bodies as process programs (include/cxxdes_b200/record/cxxdes/cxxdes.hpp).

tests/native/recorded_models.hpp is ONE model source written on the reference's public API (CXXDES_SIMULATION,
coroutine<>, subroutine<>, timeout(t, priority), all_of / any_of / && / ||, sequential / the comma operator, async,
_Co_with, sync::queue / resource / mutex, .priority(), .return_latency(), env.time_unit / time_precision / timeout).  It is compiled twice: against the reference's headers and
run by the unmodified reference engine, and against the recording <cxxdes/cxxdes.hpp>, which runs every body once on the
host and emits the model as a process program.  The program, interpreted by the CPU oracle (here) and by the CUDA
kernels (-m gpu), must dispatch the tokens the reference engine dispatched for the same source."""
import ctypes as C
import json
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracles
import libcxxdes_b200 as L
from libcxxdes_b200 import abi

ROOT = Path(__file__).resolve().parent.parent
NATIVE = ROOT / "tests" / "native"
REFERENCE = Path("/root/reference/include")
UNTIL = {"clocks": 20}
SEED, FIRST = 7, 3           # what record_reference.cpp gives cxxdes_b200::rv_context


@pytest.fixture(scope="module")
def recorded(tmp_path_factory):
    exe = tmp_path_factory.mktemp("record") / "record_dump"
    subprocess.run(["g++", "-std=c++20", "-O1", "-Wall", "-Werror", "-DCXXDES_B200_RECORDING",
                    f"-I{ROOT / 'include' / 'cxxdes_b200' / 'record'}", f"-I{ROOT / 'include'}", f"-I{NATIVE}",
                    str(NATIVE / "record_dump.cpp"), "-o", str(exe)], check=True)
    return json.loads(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)


@pytest.fixture(scope="module")
def reference_runs(tmp_path_factory):
    if not REFERENCE.is_dir():
        pytest.skip("no /root/reference on this machine")
    import torch
    torch_inc = Path(torch.__file__).parent / "include"
    fmt = ["-DFMT_HEADER_ONLY", f"-I{torch_inc}"] if (torch_inc / "fmt" / "core.h").exists() else [f"-I{ROOT / 'oracle' / 'ref' / 'fmt_stub'}"]
    exe = tmp_path_factory.mktemp("record") / "record_reference"
    subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-mfma", *fmt, f"-I{REFERENCE}",
                    f"-I{ROOT / 'oracle' / 'ref'}", f"-I{ROOT / 'include'}", f"-I{NATIVE}",
                    str(NATIVE / "record_reference.cpp"), "-o", str(exe)], check=True)
    return json.loads(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)


def program_params(tables):
    """cxxdes_b200_program from the dumped tables (keeps the arrays alive on the returned struct)."""
    ops = (abi.Op * len(tables["ops"]))(*[abi.Op(*o) for o in tables["ops"]])
    procs = (abi.Process * len(tables["procs"]))(*[abi.Process(*p) for p in tables["procs"]])
    roots = (C.c_uint32 * len(tables["roots"]))(*tables["roots"])
    qmax = (C.c_uint64 * max(1, len(tables["queue_max"])))(*tables["queue_max"])
    sinit = (C.c_uint64 * max(1, len(tables["sems"])))(*[s[0] for s in tables["sems"]])
    smax = (C.c_uint64 * max(1, len(tables["sems"])))(*[s[1] for s in tables["sems"]])
    rates = (C.c_double * max(1, len(tables["rates"])))(*tables["rates"])
    pp = abi.ProgramParams(ops, procs, roots, qmax, sinit, smax, rates, len(ops), len(procs), len(roots),
                           len(tables["queue_max"]), len(tables["sems"]), tables["n_mutexes"], tables["n_events"],
                           len(tables["rates"]), 0, 0)
    pp._keep = (ops, procs, roots, qmax, sinit, smax, rates)
    return pp


def clock_of(tables):
    (ut, uu), (pt, pu) = tables["clock"]
    return abi.Clock.of((ut, uu), (pt, pu)), ((ut, uu), (pt, pu))


def test_recorded_programs_pass_the_program_check(recorded, cuda_lib):
    for name, tables in recorded.items():
        pp = program_params(tables)
        assert cuda_lib.cxxdes_b200_check_program(C.byref(pp)) == abi.OK, (name, cuda_lib.cxxdes_b200_last_error())


def test_loops_are_folded_back(recorded):
    """400 turns of `co_await q.put(..); co_await env.timeout(lambda());` are one LOOP, not 800 operations; a body
    that never ends is a loop of 2^30 turns."""
    pc = recorded["producer_consumer"]["ops"]
    assert len(pc) == 15 and [abi.OP_LOOP, 0, 8, 400] in pc and [abi.OP_LOOP, 0, 13, 399] in pc
    assert sum(1 for o in recorded["clocks"]["ops"] if o[0] == abi.OP_LOOP and o[3] == 1 << 30) == 3


@pytest.mark.parametrize("name", ["resource", "mutex", "producer_consumer", "clocks", "mixed", "co_with", "sequential"])
def test_recorded_program_equals_the_reference_engine_running_the_source(recorded, reference_runs, port, name):
    tables = recorded[name]
    pp = program_params(tables)
    clock, _ = clock_of(tables)
    until = UNTIL.get(name, -1)
    runs = reference_runs[name]
    got = port.run(abi.MODEL_PROGRAM, pp, clock, SEED, FIRST, len(runs), threads=2, until=until)
    for r, want in enumerate(runs):
        assert int(got.n_events[r]) == want["n_events"] and int(got.end_time[r]) == want["end_time"], (name, r)
        total, rows = port.trace(abi.MODEL_PROGRAM, pp, clock, SEED, FIRST + r, max_events=4000, until=until)
        assert total == want["n_events"]
        assert [list(e[:3]) for e in rows] == want["rows"], f"{name}, replication {r}: (time, priority, kind) differ"


def test_recorded_producer_consumer_is_the_mm1_model(recorded, port):
    """The recorded examples/producer_consumer.cpp dispatches exactly the tokens of the M/M/1 model (same ordinals,
    same Philox streams): event counts, end times and trace digests are equal."""
    tables = recorded["producer_consumer"]
    clock, _ = clock_of(tables)
    a = port.run(abi.MODEL_PROGRAM, program_params(tables), clock, 11, 100, 64, threads=4)
    b = port.run(abi.MODEL_MM1, abi.MM1Params(5.0, 10.0, 400), clock, 11, 100, 64, threads=4)
    assert np.array_equal(a.n_events, b.n_events) and np.array_equal(a.end_time, b.end_time)
    assert np.array_equal(a.trace_hash, b.trace_hash)


class _RawProgram:
    """What libcxxdes_b200.Ensemble needs from a program object."""

    def __init__(self, tables):
        self._pp = program_params(tables)

    def to_params(self):
        return self._pp


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["resource", "mutex", "producer_consumer", "clocks", "mixed", "co_with", "sequential"])
def test_gpu_runs_the_recorded_programs(recorded, port, name):
    tables = recorded[name]
    clock, pairs = clock_of(tables)
    until = UNTIL.get(name, -1)
    n = 300
    ens = L.Ensemble(_RawProgram(tables), n, seed=SEED, first_replication=FIRST)
    ens.time_unit(*pairs[0]).time_precision(*pairs[1])
    got = (ens.run() if until < 0 else ens.run_until(until)).fetch(strict=False)
    ens.close()
    want = port.run(abi.MODEL_PROGRAM, program_params(tables), clock, SEED, FIRST, n, threads=os.cpu_count(), until=until)
    assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
    assert np.array_equal(got.n_events, want.n_events) and np.array_equal(got.end_time, want.end_time)
    assert np.array_equal(got.trace_hash, want.trace_hash)


def _build_example(tmp_path):
    exe = tmp_path / "recorded_resource"
    lib_dir = ROOT / "libcxxdes_b200" / "_lib"
    subprocess.run(["g++", "-std=c++20", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include' / 'cxxdes_b200' / 'record'}",
                    f"-I{ROOT / 'include'}", str(ROOT / "examples" / "recorded_resource.cpp"), f"-L{lib_dir}", "-lcxxdes_b200",
                    f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    return exe


def test_recorded_example_builds_and_fails_loudly_without_gpu(built, tmp_path):
    """examples/recorded_resource.cpp = the reference's examples/resource.cpp with its coroutine bodies untouched; there
    is no CPU path behind sim.ensemble(): without a GPU it reports the C ABI's error."""
    import torch
    exe = _build_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; covered by the gpu test")
    r = subprocess.run([str(exe), "8"], capture_output=True, text=True)
    assert r.returncode == 2 and "cxxdes_b200:" in r.stderr


@pytest.mark.gpu
def test_recorded_example_runs_on_gpu(built, tmp_path):
    exe = _build_example(tmp_path)
    r = subprocess.run([str(exe), "20000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "now() of replication 0: 12" in r.stdout and "events 320000" in r.stdout
