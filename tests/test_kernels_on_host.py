"""The `-m gpu` parity tests' LOGIC where no GPU exists: the library's own .cu sources, compiled by g++ against the CUDA
stand-in of tests/native/cuda_on_host (every CUDA thread a fibre; votes, shuffles, __syncwarp and __syncthreads real
rendez-vous points; shared memory an arena; explicit ld/st.shared PTX turned into the same accesses), run the GPU test
files in a child pytest whose CXXDES_B200_LIBRARY names that build.  Every test goes through the C ABI exactly as on the
device and compares bit for bit with the same golden vectors and oracles.

This is TEST INFRASTRUCTURE, not a CPU path of the product: the emulated library lives under tests/native/_build, nothing
in libcxxdes_b200/ refers to it, and it says nothing about speed, memory ordering between warps or resource limits --
those are what the B200 run of the same tests is for.  What it does catch before a GPU is available: a state machine,
heap, hand-over or saved-state bug in any kernel (all of them compile here: the three M/M/1 kernels, ALOHA on 8- and
4-byte tokens and with cooperative heaps, M/M/c, the memory hierarchy, both interpreters, the reductions).

The C++ hosts of examples/ (the host mirror ensemble.hpp, the program builder, the recording <cxxdes/cxxdes.hpp>) link
`-lcxxdes_b200`; the child's LD_LIBRARY_PATH names a directory where that name is the emulated library, so the same
binaries run here too.

Left out (DESELECT): tests sized for the device (minutes here), tests whose point is concurrency between two kernels
(streams run in host order here, so the M/M/1 helper never overlaps the first pass), and the multi-GPU NCCL example."""
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "native" / "cuda_on_host"))

SIZED_FOR_THE_DEVICE = "sized for the device"
NEEDS_CONCURRENT_KERNELS = "needs two kernels running at once"
NEEDS_GPUS_AND_NCCL = "one thread per GPU and an NCCL communicator"
DESELECT = {
    "test_mm1_gpu.py": {
        "test_headline_config_full_size_against_the_oracle": SIZED_FOR_THE_DEVICE,
        "test_full_size_properties": SIZED_FOR_THE_DEVICE,
        "test_pieces_on_an_ensemble_larger_than_the_grid": SIZED_FOR_THE_DEVICE,
        "test_headline_second_pass_runs_next_to_the_first": NEEDS_CONCURRENT_KERNELS,
        "test_second_pass_list_longer_than_the_helper": NEEDS_CONCURRENT_KERNELS,
    },
    "test_models_gpu.py": {
        "test_aloha_full_size_properties": SIZED_FOR_THE_DEVICE,
        "test_archsim_full_size_properties": SIZED_FOR_THE_DEVICE,
        "test_mmc_conservation": SIZED_FOR_THE_DEVICE,
        "test_pieces_on_an_ensemble_larger_than_the_grid": SIZED_FOR_THE_DEVICE,
        "test_gpu_matches_oracle[aloha-fields0-2000]": SIZED_FOR_THE_DEVICE,
    },
    "test_narrow_tokens_gpu.py": {
        "test_mmc_token_widths_agree_at_full_length": SIZED_FOR_THE_DEVICE,
        "test_token_widths_agree_on_a_large_sweep": SIZED_FOR_THE_DEVICE,
        "test_archsim_layouts_agree_at_size_and_in_pieces": SIZED_FOR_THE_DEVICE,
        "test_wide_tokens_match_oracle[fields0-2000]": SIZED_FOR_THE_DEVICE,
    },
    "test_coop_heap.py": {      # groups of 8 lanes meet a dozen times per event: ~100 us per event here
        "test_gpu_cooperative_aloha_matches_oracle[fields0-2000]": SIZED_FOR_THE_DEVICE,
        "test_gpu_cooperative_aloha_matches_oracle[fields1-512]": SIZED_FOR_THE_DEVICE,
        "test_gpu_cooperative_aloha_matches_oracle[fields2-1024]": SIZED_FOR_THE_DEVICE,
        "test_gpu_cooperative_aloha_matches_reference_golden[0]": SIZED_FOR_THE_DEVICE,
        "test_gpu_cooperative_aloha_equals_thread_per_replication_at_size": SIZED_FOR_THE_DEVICE,
    },
    "test_host_cpp.py": {"test_cpp_multi_gpu_example_allreduces_over_nccl": NEEDS_GPUS_AND_NCCL},
}
# test file -> the fewest tests that must have passed there (a deselection typo or a collection error must not pass as
# "nothing failed")
GPU_TEST_FILES = {
    "test_mm1_gpu.py": 59, "test_mm1_fuzz.py": 40, "test_models_gpu.py": 43, "test_models_fuzz.py": 40,
    "test_narrow_tokens_gpu.py": 56, "test_programs.py": 34, "test_programs_fuzz.py": 120, "test_trace_gpu.py": 23,
    "test_coop_heap.py": 9, "test_record.py": 8, "test_host_cpp.py": 4, "test_written_after_the_last_gpu_run.py": 281,   # (+ 5 that need the reference's examples, skipped where /root/reference is absent)
}


@pytest.fixture(scope="module")
def emulated_library(built):
    import build as cuda_on_host
    return cuda_on_host.build()


# The order in which the lanes of a warp run between two rendez-vous is undefined on the device (independent thread
# scheduling).  The emulation runs them lane 0 first, and once more in a pseudo-random permutation per pass: code that
# reads what another lane wrote without a __syncwarp in between gives different results in the two.
# The second run also gives the library the B200's 148 SMs, i.e. the grids, strides and buffer sizes of the real device
# (the first: 2 SMs, where a shard is soon larger than the grid and lanes claim second replications).
@pytest.fixture(scope="module", params=[("lane 0 first", "2"), ("random:20261020", "148")], ids=["lane 0 first, 2 SMs", "random lane order, 148 SMs"])
def emulated_run(emulated_library, request):
    """ONE child pytest over all GPU test files (interpreter and torch start-up paid once per worker), outcome per test."""
    env = dict(os.environ, CXXDES_B200_LIBRARY=str(emulated_library), CUDA_ON_HOST_SMS=request.param[1], CUDA_ON_HOST_ORDER=request.param[0],
               LD_LIBRARY_PATH=str(Path(emulated_library).parent / "lib"))
    cmd = [sys.executable, "-m", "pytest", *[f"tests/{f}" for f in sorted(GPU_TEST_FILES)], "-m", "gpu", "-q", "-rA",
           "-p", "no:cacheprovider", "-n", str(min(6, os.cpu_count() or 1)), "--timeout", "600"]
    for test_file, names in DESELECT.items():
        for name in names:
            cmd += ["--deselect", f"tests/{test_file}::{name}"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True)
    outcome = {}
    for line in r.stdout.splitlines():
        m = re.match(r"(PASSED|FAILED|ERROR) tests/(test_\w+\.py)::(\S+)", line)
        if m:
            outcome.setdefault(m.group(2), {})[m.group(3)] = m.group(1)
    return r, outcome


@pytest.mark.parametrize("test_file", sorted(GPU_TEST_FILES))
def test_gpu_tests_pass_on_the_emulated_device(emulated_run, test_file):
    r, outcome = emulated_run
    mine = outcome.get(test_file, {})
    bad = sorted(name for name, what in mine.items() if what != "PASSED")
    if bad:
        detail = re.findall(r"_{5,} (?:%s)[^\n]* _{5,}\n.*?(?=\n_{5,} |\n={5,} )" % "|".join(re.escape(b.split("[")[0]) for b in bad), r.stdout, re.S)
        pytest.fail(f"{test_file} on the emulated device: {bad}\n" + "\n".join(detail)[-6000:])
    assert len(mine) >= GPU_TEST_FILES[test_file], (len(mine), r.stdout[-2000:] + r.stderr[-2000:])


def test_every_gpu_test_file_is_listed():
    """A new -m gpu test file must be added above (or it silently never runs on the emulated device)."""
    with_gpu_tests = {p.name for p in (ROOT / "tests").glob("test_*.py")
                      if re.search(r"pytest\.mark\.gpu|pytestmark\s*=\s*pytest\.mark\.gpu", p.read_text()) and p.name != Path(__file__).name}
    assert with_gpu_tests == set(GPU_TEST_FILES)


def test_the_emulated_library_is_not_a_product_path(emulated_library):
    """Nothing under libcxxdes_b200/ nor bench.py knows the emulated build, and it does not sit where the product looks."""
    product = [p for p in (ROOT / "libcxxdes_b200").rglob("*") if p.is_file() and p.suffix in (".py", ".cu", ".cuh", ".h")]
    for p in product + [ROOT / "bench.py"]:
        assert "cuda_on_host" not in p.read_text(errors="ignore"), p
    assert ROOT / "tests" in Path(emulated_library).parents
    assert not (ROOT / "libcxxdes_b200" / "_lib" / Path(emulated_library).name).exists()
