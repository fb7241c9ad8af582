"""The host C++ layer (include/cxxdes_b200/ensemble.hpp) compiles as a plain g++ translation unit,
links against the C-ABI library, and reports errors the way the reference does (exceptions)."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def build_example(tmp_path):
    exe = tmp_path / "mm1_ensemble"
    lib_dir = ROOT / "libcxxdes_b200" / "_lib"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}",
                    str(ROOT / "examples" / "mm1_ensemble.cpp"), f"-L{lib_dir}", "-lcxxdes_b200",
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
    for lam, mu, rho, avg, _ in rows:
        lam, mu, avg = float(lam), float(mu), float(avg)
        if lam <= 7.5:   # heavier loads have not reached steady state within 2000 packets
            assert abs(avg - 1.0 / (mu - lam)) < 0.15 / (mu - lam)   # M/M/1 sojourn time 1/(mu-lambda)
        else:
            assert avg > 0.3
