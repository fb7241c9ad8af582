import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(autouse=True)
def _only_what_has_run_on_a_b200(request, monkeypatch):
    """The last passes of M/M/c and of the interpreter (heaps and wait lists in global memory) were written after the
    round's last B200 run.  The GPU tests that HAVE run on the device keep running exactly the launches they ran there:
    CXXDES_B200_DEEP_BYTES=0 (read at create) leaves those passes out.  tests/test_written_after_the_last_gpu_run.py --
    which sorts last -- runs everything with the library's defaults."""
    if (request.node.get_closest_marker("gpu") and request.node.fspath.basename != "test_written_after_the_last_gpu_run.py"
            and "CXXDES_B200_DEEP_BYTES" not in os.environ):
        monkeypatch.setenv("CXXDES_B200_DEEP_BYTES", "0")


@pytest.fixture(scope="session")
def built():
    """Everything compiled (CUDA library + CPU checkers); cheap when already up to date."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def port(built):
    import oracles
    return oracles.port_oracle()


@pytest.fixture(scope="session")
def ref(built):
    """The reference-engine oracle; tests that need it skip where it was never built."""
    import oracles
    o = oracles.ref_oracle()
    if o is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    return o


@pytest.fixture(scope="session")
def cuda_lib(built):
    import libcxxdes_b200 as L
    return L.load_library()


def host_threads():
    return max(1, os.cpu_count() or 1)
