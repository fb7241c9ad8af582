"""The reference's OWN example programs, source files unchanged, through the recording <cxxdes/cxxdes.hpp>.

Each file of /root/reference/examples listed below is compiled twice where it lies (main() renamed where the harness
makes the simulation object; kept where the example drives an `environment` by hand):
  * against the unmodified reference headers and run by the reference engine (tests/native/example_reference.cpp,
    observed through env.next_event() / env.step());
  * against include/cxxdes_b200/record (tests/native/example_recorder.cpp): its coroutine bodies, run once on the host,
    come out as a process program.
The program on the CPU interpreter must dispatch the tokens -- (time, priority, kind), event for event -- that the
reference engine dispatched for the same source: `_Co_with` over a mutex and a resource, bounded queues, semaphores,
nested compositions, `sequential` and the comma operator with compositions inside, recursion 21 processes deep,
`all_of.range` / `any_of.range`, `evt.wait() || timeout(t)`, return values handed from process to process,
`this_coroutine()`.  (The CUDA interpreters run such programs bit-equal to the CPU interpreter: tests/test_programs*.py,
tests/test_record.py, and the -m gpu test in tests/test_written_after_the_last_gpu_run.py.)  Not in the list: the four config models (hand-
written kernels), C++ coroutine experiments that do not use the library (subroutine, symmetric, stack, time), examples
that need exceptions through tokens, user-defined awaitables or as_coroutine() -- and capture_return.cpp, which crashes
on the reference itself here."""
import json
import os
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
    # time literals (`10_x`) and lazy_timeout; a coroutine handle copied into two processes that both await it, awaited
    # again after it has returned, and left running behind `p || delay(1)`
    ("lazy_timeout", "instant_test", None), ("compositions_lvalueref", "async_example", None),
    # examples that drive an `environment` by hand -- env.bind(...) several times (root processes), env.run() / run_for(t):
    # class None = the example's own main() runs (-DEXAMPLE_MAIN)
    ("event", None, None), ("clocks", None, None), ("get_environment", None, None), ("bug", None, None),
    ("process_rewrite", None, None),
]
# class-shaped examples whose main() passes constructor arguments / runs for a while: (EXAMPLE_ARGS, EXAMPLE_UNTIL)
# several simulations in one file: (label, class, constructor arguments or None, horizon of its run_for or None)
VARIANTS = {"interrupt_test": [("0", "test", "0", 100), ("1", "test", "1", 100), ("2", "test", "2", 100)],
            # a coroutine member awaited by reference from two places; subroutines three deep; a subroutine stopped by
            # run_for in the middle; `foo().as_coroutine() && bar()` (its fifth class recurses ten million deep: left out)
            "pitfall": [("pitfall", "pitfall", None, None), ("subroutines", "subroutines", None, None),
                        ("subroutines2", "subroutines2", None, 500), ("subroutines3", "subroutines3", None, 500)]}
for _name, _variants in VARIANTS.items():
    for _label, _sim, _args, _until in _variants:
        EXAMPLES.append((f"{_name}:{_label}", _sim, None))


def _build_and_run(tmp, name, sim, on_device=False):
    """on_device: without -DCXXDES_B200_RECORD_ONLY -- the example's own env.run*() runs one replication through the C ABI."""
    import torch
    torch_inc = Path(torch.__file__).parent / "include"
    fmt = ["-DFMT_HEADER_ONLY", f"-I{torch_inc}"] if (torch_inc / "fmt" / "core.h").exists() else [f"-I{ROOT / 'oracle' / 'ref' / 'fmt_stub'}"]
    file_name, _, variant = name.partition(":")
    src = REFERENCE / "examples" / f"{file_name}.cpp"
    defs = [f'-DEXAMPLE_FILE="{src}"']
    if sim is None:
        defs.append("-DEXAMPLE_MAIN")
    else:
        defs.append(f"-DEXAMPLE_SIM={sim}")
    if variant:
        _, _, args, until = next(v for v in VARIANTS[file_name] if v[0] == variant)
        defs += ([f"-DEXAMPLE_ARGS={args}"] if args is not None else []) + ([f"-DEXAMPLE_UNTIL={until}"] if until is not None else [])
    lib_dir = ROOT / "libcxxdes_b200" / "_lib"
    tag = name.replace(":", "_")
    rec, ref = tmp / f"{tag}_rec", tmp / f"{tag}_ref"
    subprocess.run(["g++", "-std=c++20", "-O1", "-w", *fmt, "-DCXXDES_B200_RECORDING", *([] if on_device else ["-DCXXDES_B200_RECORD_ONLY"]),
                    f"-I{ROOT / 'include' / 'cxxdes_b200' / 'record'}",
                    f"-I{ROOT / 'include'}", f"-I{NATIVE}", *defs, str(NATIVE / "example_recorder.cpp"), f"-L{lib_dir}",
                    "-lcxxdes_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(rec)], check=True)
    subprocess.run(["g++", "-std=c++20", "-O1", "-w", "-ffp-contract=off", "-mfma", *fmt, f"-I{REFERENCE / 'include'}",
                    f"-I{ROOT / 'oracle' / 'ref'}", f"-I{ROOT / 'include'}", f"-I{NATIVE}", *defs,
                    str(NATIVE / "example_reference.cpp"), "-o", str(ref)], check=True)
    import os
    for exe in (rec, ref):   # (class shape: the file is argv[1]; main shape: $EXAMPLE_OUT, written at exit)
        subprocess.run([str(exe), str(exe) + ".json"], check=True, capture_output=True, env=dict(os.environ, EXAMPLE_OUT=str(exe) + ".json"))
    recorded = json.loads(Path(str(rec) + ".json").read_text())
    if on_device:
        return recorded["now"], json.loads(Path(str(ref) + ".json").read_text())
    return recorded["program"], json.loads(Path(str(ref) + ".json").read_text()), recorded["until"]


@pytest.fixture(scope="module")
def examples(built, tmp_path_factory):
    if not (REFERENCE / "examples").is_dir():
        pytest.skip("no /root/reference on this machine")
    tmp = tmp_path_factory.mktemp("examples")
    with ThreadPoolExecutor(max_workers=8) as pool:
        jobs = {name: pool.submit(_build_and_run, tmp, name, sim) for name, sim, _ in EXAMPLES}
        return {name: job for name, job in jobs.items()}     # (.result() in the test: one example's failure is its own)


@pytest.mark.parametrize("name,sim,n_events", EXAMPLES)
def test_reference_example_recorded_unchanged_equals_the_reference_engine(examples, port, cuda_lib, name, sim, n_events):
    import ctypes as C
    tables, want, until = examples[name].result()
    pp = T.program_params(tables)
    assert cuda_lib.cxxdes_b200_check_program(C.byref(pp)) == abi.OK, cuda_lib.cxxdes_b200_last_error()
    clock, _ = T.clock_of(tables)
    assert want["n_events"] > 0 and (n_events is None or want["n_events"] == n_events)   # (the reference engine ran the example)
    got = port.run(abi.MODEL_PROGRAM, pp, clock, T.SEED, T.FIRST, 1, threads=1, until=until)
    assert int(got.status[0]) & ~abi.REP_UNFINISHED == 0
    assert int(got.n_events[0]) == want["n_events"] and int(got.end_time[0]) == want["end_time"]
    total, rows = port.trace(abi.MODEL_PROGRAM, pp, clock, T.SEED, T.FIRST, max_events=20000, until=until)
    assert total == want["n_events"]
    assert [list(e[:3]) for e in rows] == want["rows"], f"{name}: (time, priority, kind) differ"


# ---- the same on the device: tests/test_written_after_the_last_gpu_run.py reads the recordings committed here ----------
# /root/reference does not travel to the GPU box: the recorded programs (tables derived from the examples, not their
# source) and what the reference engine dispatched for each are committed as tests/golden/recorded_examples.json by
# `python tests/test_record_examples.py` (the generator: the fixture above, written out).
GOLDEN = ROOT / "tests" / "golden" / "recorded_examples.json"


def _golden():
    return json.loads(GOLDEN.read_text())["examples"]


def test_committed_recordings_are_current(examples):
    """The committed file is what recording the examples here gives (same programs, same reference rows)."""
    committed = _golden()
    assert sorted(committed) == sorted(name for name, _, _ in EXAMPLES)
    for name, _, _ in EXAMPLES:
        tables, want, until = examples[name].result()
        assert committed[name] == {"program": tables, "reference": want, "until": until}, name


if __name__ == "__main__":      # regenerate tests/golden/recorded_examples.json (needs /root/reference)
    import sys
    import tempfile
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as g
    g.build()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, sim, _ in EXAMPLES:
            tables, want, until = _build_and_run(Path(tmp), name, sim)
            out[name] = {"program": tables, "reference": want, "until": until}
    GOLDEN.write_text(json.dumps({"about": "reference examples recorded by include/cxxdes_b200/record (program tables) and what the "
                                           "unmodified reference engine dispatched for the same source files (rows: time, priority, kind); "
                                           "generator: python tests/test_record_examples.py", "examples": out}, indent=None, separators=(",", ":")) + "\n")
    print(f"wrote {GOLDEN} ({GOLDEN.stat().st_size} bytes, {len(out)} examples)")
