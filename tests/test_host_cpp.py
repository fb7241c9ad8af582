"""The host C++ layer (include/cxxdes_b200/ensemble.hpp) compiles as a plain g++ translation unit,
links against the C-ABI library, and reports errors the way the reference does (exceptions)."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def build_example(tmp_path, name="mm1_ensemble"):
    exe = tmp_path / name
    lib_dir = ROOT / "libcxxdes_b200" / "_lib"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}",
                    str(ROOT / "examples" / f"{name}.cpp"), f"-L{lib_dir}", "-lcxxdes_b200",
                    f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    return exe


def test_cpp_example_builds_and_fails_loudly_without_gpu(built, tmp_path):
    import torch
    exe = build_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; covered by the gpu test")
    r = subprocess.run([str(exe), "8", "50"], capture_output=True, text=True)
    assert r.returncode == 2
    assert "cxxdes_b200:" in r.stderr          # std::runtime_error carrying the C-ABI error text


@pytest.mark.gpu
def test_cpp_example_runs_on_gpu(built, tmp_path):
    exe = build_example(tmp_path)
    r = subprocess.run([str(exe), "2048", "2000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = [line.split(",") for line in r.stdout.strip().splitlines()[1:]]
    assert len(rows) == 5
    for lam, mu, rho, avg, ci, _ in rows:
        lam, mu, avg = float(lam), float(mu), float(avg)
        assert 0.0 < float(ci) < 0.1 * avg
        if lam <= 7.5:   # heavier loads have not reached steady state within 2000 packets
            assert abs(avg - 1.0 / (mu - lam)) < 0.15 / (mu - lam)   # M/M/1 sojourn time 1/(mu-lambda)
        else:
            assert avg > 0.3


# ------------------------------------------------------------------ host C++ builder for process programs
def _python_tables(prog):
    prog.to_params()
    return {"ops": [list(o) for o in prog._ops],
            "procs": [[p["entry"], p["ordinal"], p["priority"], p["latency"], p["return_latency"], p["return_priority"]]
                      for p in prog._procs],
            "roots": list(prog._roots), "queue_max": list(prog._queue_max), "sems": [list(x) for x in prog._sems],
            "rates": [r[step] if isinstance(r, list) else r for step in range(max(1, prog._sweep_steps)) for r in prog._rates],
            "sweep_steps": prog._sweep_steps, "n_mutexes": prog._n_mutexes, "n_events": prog._n_events}


def test_cpp_program_builder_emits_the_pinned_programs(tmp_path):
    """include/cxxdes_b200/program.hpp against the Python builder: the same models (M/M/1 and thirteen of the
    reference's own test/example scenarios, whose Python programs are pinned on the reference engine by
    tests/test_programs.py) must lower to identical cxxdes_b200_program tables.  Built as C++17 and C++20
    (the reference's users compile as C++20, where `co_await` is a keyword)."""
    import json
    import kat_programs as K
    src = ROOT / "tests" / "native" / "program_builder_dump.cpp"
    subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(src)], check=True)
    exe = tmp_path / "program_builder_dump"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(src),
                    "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0           # also: the three misuse cases threw
    got = json.loads(r.stdout)
    assert got.pop("mm1") == _python_tables(K.mm1_program(5.0, 10.0, 300))
    assert got.pop("mm1k") == _python_tables(K.mm1k_program(8.0, 10.0, 400, 4))
    assert got.pop("mm1_sweep") == _python_tables(K.mm1_program([1.0, 3.0, 5.0, 7.0, 9.0], 10.0, 300))
    assert sorted(got) == sorted(str(k) for k in (1, 2, 4, 5, 6, 7, 9, 10, 11, 14, 15, 16, 17))
    for k, tables in got.items():
        assert tables == _python_tables(K.SCENARIOS[int(k)]()[0]), K.SCENARIOS[int(k)].__name__


@pytest.mark.gpu
def test_cpp_program_example_matches_the_fused_kernel(built, tmp_path):
    exe = build_example(tmp_path, "program_mm1")
    r = subprocess.run([str(exe), "4096", "400"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "differing replications 0" in r.stdout


@pytest.mark.gpu
def test_cpp_aloha_sweep_example(built, tmp_path):
    """examples/aloha_sweep.cpp prints the table of the reference's examples/aloha.cpp from one ensemble run."""
    import math
    exe = build_example(tmp_path, "aloha_sweep")
    r = subprocess.run([str(exe), "1000", "16", "200"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0].startswith("G [Erlang], S [Erlang]") and len(lines) == 51
    for line in lines[1:]:
        G, S, ci = (float(x) for x in line.split(", ")[:3])
        assert abs(S - G * math.exp(-2 * G)) < 0.04 and ci < 0.05      # pure ALOHA throughput


def build_multi_gpu_example(tmp_path):
    exe = tmp_path / "mm1_multi_gpu"
    lib_dir = ROOT / "libcxxdes_b200" / "_lib"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-pthread", f"-I{ROOT / 'include'}",
                    "-I/usr/local/cuda/include", str(ROOT / "examples" / "mm1_multi_gpu.cpp"), f"-L{lib_dir}",
                    "-lcxxdes_b200", "-lnccl", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    return exe


def test_reduce_allgpus_rejects_misuse_without_touching_nccl(cuda_lib):
    """The collective entry point validates its arguments before it looks for NCCL (the library does not link it)."""
    import libcxxdes_b200 as L
    assert cuda_lib.cxxdes_b200_reduce_allgpus(None, None, None) == L.abi.ERR_INVALID
    assert b"handle is NULL" in cuda_lib.cxxdes_b200_last_error()


@pytest.mark.gpu
def test_cpp_multi_gpu_example_allreduces_over_nccl(built, tmp_path):
    """One thread per GPU, shards by first_replication, ensemble::reduce_allgpus over an ncclComm_t: the totals are
    the sums of the shard accumulators on every rank -- and the same digest sum whatever the number of GPUs."""
    import torch
    exe = build_multi_gpu_example(tmp_path)
    n = torch.cuda.device_count()
    outs = []
    for gpus, per_gpu in ((n, 4096 // n * 2), (1, 4096 // n * 2 * n)):   # same global replications, n shards vs 1
        r = subprocess.run([str(exe), str(per_gpu), "300", str(gpus)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "inconsistent ranks 0" in r.stdout
        outs.append(r.stdout)
    field = lambda s, key: s.split(key)[1].split(",")[0].strip()
    assert field(outs[0], "digest sum") == field(outs[1], "digest sum")
    assert field(outs[0], "events") == field(outs[1], "events")


@pytest.mark.gpu
def test_cpp_loss_system_example_matches_theory(built, tmp_path):
    """examples/mm1k_loss.cpp: an `if / else` between awaits and a load sweep inside one ensemble, from the C++
    builder; the dropped fraction follows the M/M/1/K blocking probability at every load point (exit code 0)."""
    exe = build_example(tmp_path, "mm1k_loss")
    r = subprocess.run([str(exe), "20000", "4000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert len(r.stdout.strip().splitlines()) == 6


def test_cpp_planned_kernel_without_a_device(built, tmp_path):
    """ensemble<Model>::planned_kernel() (cxxdes_b200_plan_kernel): the first-pass kernel of a configuration, asked from
    the C++ mirror before anything is created."""
    exe = tmp_path / "plan_kernel_host"
    lib_dir = ROOT / "libcxxdes_b200" / "_lib"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}",
                    str(ROOT / "tests" / "native" / "plan_kernel_host.cpp"), f"-L{lib_dir}", "-lcxxdes_b200",
                    f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "aloha aloha_compact_kernel<false>"
    assert lines[1] == "mmc mmc_fast_kernel<false,wide>"            # 1023 customers: one ordinal too many for 4-byte tokens
    assert lines[2] == "archsim archsim_kernel<k requesters,narrow>"
    assert lines[3].startswith("mm1 throws: cxxdes_b200:")
