"""The reference's OWN example programs, source files unchanged, through the recording <cxxdes/cxxdes.hpp>.

Each file of /root/reference/examples listed below is compiled twice where it lies -- only its main() is renamed:
  * against the unmodified reference headers and run by the reference engine (tests/native/example_reference.cpp,
    observed through env.next_event() / env.step());
  * against include/cxxdes_b200/record (tests/native/example_recorder.cpp): its coroutine bodies, run once on the host,
    come out as a process program.
The program on the CPU interpreter must dispatch the tokens -- (time, priority, kind), event for event -- that the
reference engine dispatched for the same source: `_Co_with` over a mutex and a resource, bounded queues, semaphores,
nested compositions, `sequential` and the comma operator with compositions inside, recursion 21 processes deep,
`all_of.range` / `any_of.range`, `evt.wait() || timeout(t)`, return values handed from process to process,
`this_coroutine()`.  (The CUDA interpreters run such programs bit-equal to the CPU interpreter: tests/test_programs*.py,
tests/test_record.py.)  Not in the list: examples that drive an `environment` by hand in main(), that use the time-
expression literals (`10_x`), copyable coroutine handles, interrupts or exceptions -- and capture_return.cpp, which
crashes on the reference itself here."""
import json
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import pytest

from libcxxdes_b200 import abi
import test_record as T

ROOT = Path(__file__).resolve().parent.parent
NATIVE = ROOT / "tests" / "native"
REFERENCE = Path("/root/reference")
EXAMPLES = [            # file, its CXXDES_SIMULATION class, events the reference engine dispatches
    ("any_of", "any_of_example", 15), ("mutex", "mutex_example", 38), ("queue", "queue_example", 11),
    ("raii_trick", "raii_trick", 29), ("resource", "resource_example", 16), ("semaphore", "queue_example", 14),
    ("return_value", "return_value_example", 9), ("return_value_and_timeout", "bug", 8),
    ("complicated", "complicated_example", 190), ("this_out_scope", "this_out_scope", 8), ("debug", "debug_example", 11),
]


def _build_and_run(tmp, name, sim):
    import torch
    torch_inc = Path(torch.__file__).parent / "include"
    fmt = ["-DFMT_HEADER_ONLY", f"-I{torch_inc}"] if (torch_inc / "fmt" / "core.h").exists() else [f"-I{ROOT / 'oracle' / 'ref' / 'fmt_stub'}"]
    src = REFERENCE / "examples" / f"{name}.cpp"
    defs = [f'-DEXAMPLE_FILE="{src}"', f"-DEXAMPLE_SIM={sim}"]
    lib_dir = ROOT / "libcxxdes_b200" / "_lib"
    rec, ref = tmp / f"{name}_rec", tmp / f"{name}_ref"
    subprocess.run(["g++", "-std=c++20", "-O1", "-w", *fmt, "-DCXXDES_B200_RECORDING", f"-I{ROOT / 'include' / 'cxxdes_b200' / 'record'}",
                    f"-I{ROOT / 'include'}", f"-I{NATIVE}", *defs, str(NATIVE / "example_recorder.cpp"), f"-L{lib_dir}",
                    "-lcxxdes_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(rec)], check=True)
    subprocess.run(["g++", "-std=c++20", "-O1", "-w", "-ffp-contract=off", "-mfma", *fmt, f"-I{REFERENCE / 'include'}",
                    f"-I{ROOT / 'oracle' / 'ref'}", f"-I{ROOT / 'include'}", f"-I{NATIVE}", *defs,
                    str(NATIVE / "example_reference.cpp"), "-o", str(ref)], check=True)
    subprocess.run([str(rec), str(rec) + ".json"], check=True, capture_output=True)
    subprocess.run([str(ref), str(ref) + ".json"], check=True, capture_output=True)
    return json.loads(Path(str(rec) + ".json").read_text())["program"], json.loads(Path(str(ref) + ".json").read_text())


@pytest.fixture(scope="module")
def examples(built, tmp_path_factory):
    if not (REFERENCE / "examples").is_dir():
        pytest.skip("no /root/reference on this machine")
    tmp = tmp_path_factory.mktemp("examples")
    with ThreadPoolExecutor(max_workers=8) as pool:
        jobs = {name: pool.submit(_build_and_run, tmp, name, sim) for name, sim, _ in EXAMPLES}
        return {name: job.result() for name, job in jobs.items()}


@pytest.mark.parametrize("name,sim,n_events", EXAMPLES)
def test_reference_example_recorded_unchanged_equals_the_reference_engine(examples, port, cuda_lib, name, sim, n_events):
    import ctypes as C
    tables, want = examples[name]
    pp = T.program_params(tables)
    assert cuda_lib.cxxdes_b200_check_program(C.byref(pp)) == abi.OK, cuda_lib.cxxdes_b200_last_error()
    clock, _ = T.clock_of(tables)
    assert want["n_events"] == n_events                       # (the reference engine ran the whole example)
    got = port.run(abi.MODEL_PROGRAM, pp, clock, T.SEED, T.FIRST, 1, threads=1, until=-1)
    assert int(got.status[0]) == 0
    assert int(got.n_events[0]) == want["n_events"] and int(got.end_time[0]) == want["end_time"]
    total, rows = port.trace(abi.MODEL_PROGRAM, pp, clock, T.SEED, T.FIRST, max_events=20000, until=-1)
    assert total == want["n_events"]
    assert [list(e[:3]) for e in rows] == want["rows"], f"{name}: (time, priority, kind) differ"
