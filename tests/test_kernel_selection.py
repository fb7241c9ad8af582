"""Which first-pass kernel a configuration takes (cxxdes_b200_plan_kernel): decided on the host from the parameters, the
clock and the flags, so the rules are pinned here without a GPU -- 4-byte tokens where the delays almost surely fit the
relative time of a token and the ordinals fit the code, the narrow layout of the memory hierarchy where the run provably
ends before tick 2^31, the 8-byte / 64-bit representations otherwise.  (tests/test_narrow_tokens_gpu.py runs both sides
of every rule against the oracle on the GPU.)"""
import pytest

import libcxxdes_b200 as L
from libcxxdes_b200 import abi

S, MS, US, NS = L.SECONDS, L.MILLISECONDS, L.MICROSECONDS, L.NANOSECONDS


def plan(params, precision=(1, MS), flags=0, monkeypatch=None):
    return L.plan_kernel(params, unit=(1, S), precision=precision, flags=flags)


def test_baseline_configs(cuda_lib):
    assert plan(abi.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0)) == "aloha_compact_kernel<false>"
    assert plan(abi.MMCParams(9.0, 1.0, 0.5, 8, 0, 1000)) == "mmc_fast_kernel<false,narrow>"
    assert plan(abi.ArchSimParams(1, 10, 16, 32, 1, 0, 1024), precision=(1, S)) == "archsim_kernel<1 requester,narrow>"
    assert plan(abi.ArchSimParams(1, 10, 16, 32, 4, 0, 256), precision=(1, S)) == "archsim_kernel<k requesters,narrow>"


def test_flags_select_the_wide_representations(cuda_lib):
    w = abi.FLAG_WIDE_TOKENS
    assert plan(abi.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0), flags=w) == "aloha_fast_kernel<false>"
    assert plan(abi.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0), flags=abi.FLAG_COOP_HEAP) == "aloha_coop_kernel"
    assert plan(abi.MMCParams(9.0, 1.0, 0.5, 8, 0, 1000), flags=w) == "mmc_fast_kernel<false,wide>"
    assert plan(abi.ArchSimParams(1, 10, 16, 32, 4, 0, 256), precision=(1, S), flags=w) == "archsim_kernel<k requesters>"


def test_aloha_waiting_times_against_24_bits(cuda_lib, monkeypatch):
    """18 mean waits (N / lambda) must stay below 2^24 - 2^22 ticks: 64 stations at 1 ms fit down to G = 0.092."""
    monkeypatch.delenv("CXXDES_B200_ALOHA_TAIL", raising=False)
    assert plan(abi.AlohaParams(64, 0, 0.092, 0.092, 1.0, 1000.0, 10)) == "aloha_compact_kernel<false>"
    assert plan(abi.AlohaParams(64, 0, 0.09, 0.09, 1.0, 1000.0, 10)) == "aloha_fast_kernel<false>"
    assert plan(abi.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0), precision=(100, US)) == "aloha_fast_kernel<false>"   # 10x finer ticks
    assert plan(abi.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0), precision=(1, NS)) == "aloha_kernel"                # beyond 32-bit ticks
    # the sweep is judged by its lowest load, in either direction
    assert plan(abi.AlohaParams(64, 50, 3.0, 0.05, 1.0, 1000.0, 0)) == "aloha_fast_kernel<false>"
    # 16-bit station counters
    assert plan(abi.AlohaParams(8, 0, 2.0, 2.0, 1.0, 1000.0, 65534)) == "aloha_compact_kernel<false>"
    assert plan(abi.AlohaParams(8, 0, 2.0, 2.0, 1.0, 1000.0, 65535)) == "aloha_fast_kernel<false>"
    monkeypatch.setenv("CXXDES_B200_ALOHA_TAIL", "3")          # the tests' switch: accept a 5 % tail
    assert plan(abi.AlohaParams(64, 0, 0.02, 0.02, 1.0, 1000.0, 10)) == "aloha_compact_kernel<false>"


def test_mmc_ordinals_and_delays_against_12_and_20_bits(cuda_lib, monkeypatch):
    monkeypatch.delenv("CXXDES_B200_MMC_TAIL", raising=False)
    assert plan(abi.MMCParams(9.0, 1.0, 0.5, 8, 0, 1022)) == "mmc_fast_kernel<false,narrow>"
    assert plan(abi.MMCParams(9.0, 1.0, 0.5, 8, 0, 1023)) == "mmc_fast_kernel<false,wide>"
    # 18 mean delays of the slower rate below 2^20 - 2^18 = 786 432 ticks
    assert plan(abi.MMCParams(9.0, 0.023, 0.5, 8, 0, 100)) == "mmc_fast_kernel<false,narrow>"
    assert plan(abi.MMCParams(9.0, 0.0228, 0.5, 8, 0, 100)) == "mmc_fast_kernel<false,wide>"
    assert plan(abi.MMCParams(0.02, 1.0, 0.5, 8, 0, 100)) == "mmc_fast_kernel<false,wide>"
    assert plan(abi.MMCParams(9.0, 1.0, 800.0, 8, 0, 100)) == "mmc_fast_kernel<false,wide>"       # the patience itself
    assert plan(abi.MMCParams(9.0, 1.0, 0.5, 8, 0, 1000), precision=(10, US)) == "mmc_fast_kernel<false,wide>"
    assert plan(abi.MMCParams(9.0, 1.0, 0.5, 8, 0, 1000), precision=(1, NS)) == "mmc_kernel"


def test_archsim_bound_on_the_end_of_the_run(cuda_lib):
    """requesters x loads x (l1 + l2 + 1) < 2^31, address space <= 65 536, fewer than 65 536 loads per requester."""
    s = (1, S)
    assert plan(abi.ArchSimParams(1, 10, 16, 32, 16, 0, 65535), precision=s) == "archsim_kernel<k requesters,narrow>"
    assert plan(abi.ArchSimParams(1, 10, 16, 32, 16, 0, 65536), precision=s) == "archsim_kernel<k requesters>"
    assert plan(abi.ArchSimParams(1, 10, 16, 65536, 2, 0, 100), precision=s) == "archsim_kernel<k requesters,narrow>"
    assert plan(abi.ArchSimParams(1, 10, 16, 65537, 2, 0, 100), precision=s) == "archsim_kernel<k requesters>"
    assert plan(abi.ArchSimParams(1000000, 1000000, 16, 32, 2, 0, 536), precision=s) == "archsim_kernel<k requesters,narrow>"
    assert plan(abi.ArchSimParams(1000000, 1000000, 16, 32, 2, 0, 537), precision=s) == "archsim_kernel<k requesters>"


def test_models_planned_on_the_device_and_bad_arguments(cuda_lib):
    with pytest.raises(L.CxxdesError) as err:
        plan(abi.MM1Params(5.0, 10.0, 10000))
    assert err.value.code == abi.ERR_STATE
    lib = L.load_library()
    import ctypes as C
    cfg = abi.Config()
    cfg.model = abi.MODEL_ALOHA
    buf = C.create_string_buffer(64)
    assert lib.cxxdes_b200_plan_kernel(C.byref(cfg), buf, 64) == abi.ERR_INVALID        # no params
    assert lib.cxxdes_b200_plan_kernel(None, buf, 64) == abi.ERR_INVALID
