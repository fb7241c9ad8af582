"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every declared symbol,
and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

import libcxxdes_b200 as L

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_are_exported(cuda_lib):
    header = (ROOT / "include" / "cxxdes_b200.h").read_text()
    declared = set(re.findall(r"\b(cxxdes_b200_[a-z_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(cuda_lib, sym), f"{sym} declared in cxxdes_b200.h but not exported"
    assert declared == set(L.EXPORTED_SYMBOLS)


def test_struct_sizes_match_header():
    assert C.sizeof(L.abi.Clock) == 24
    assert C.sizeof(L.MM1Params) == 24
    assert C.sizeof(L.AlohaParams) == 32
    assert C.sizeof(L.MMCParams) == 40
    assert C.sizeof(L.ArchSimParams) == 40
    assert C.sizeof(L.abi.Columns) == 8 * (4 + L.abi.N_F64 + L.abi.N_I64)
    assert C.sizeof(L.abi.Accumulators) == 8 * (5 + L.abi.N_I64 + 2 * L.abi.N_F64)
    assert C.sizeof(L.abi.Config) == 8 + 8 + 8 + 8 + 24 + 8 + 8 + 8 + 8


def test_version(cuda_lib):
    abi_v, sm = C.c_int(), C.c_int()
    assert cuda_lib.cxxdes_b200_version(C.byref(abi_v), C.byref(sm)) == 0
    assert abi_v.value == 1 and sm.value == 100


def test_invalid_arguments_are_rejected(cuda_lib):
    h = C.c_void_p()
    assert cuda_lib.cxxdes_b200_create(None, C.byref(h)) == L.abi.ERR_INVALID
    cfg = L.abi.Config()
    cfg.model = 99
    assert cuda_lib.cxxdes_b200_create(C.byref(cfg), C.byref(h)) == L.abi.ERR_INVALID
    assert b"unknown model" in cuda_lib.cxxdes_b200_last_error()
    assert cuda_lib.cxxdes_b200_run(None) == L.abi.ERR_INVALID


def test_no_cpu_fallback_without_gpu(cuda_lib):
    """On a machine without a CUDA device, constructing an ensemble must raise, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ens = L.Ensemble(L.MM1Params(5.0, 10.0, 10), 4)
    with pytest.raises(L.CxxdesError) as ei:
        ens.run()
    assert ei.value.code == L.abi.ERR_CUDA


def test_reconfigure_after_use_raises():
    """environment.ipp:43-65: time_unit/time_precision throw once the environment is used."""
    ens = L.Ensemble(L.MM1Params(5.0, 10.0, 10), 4)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    ens._handle = C.c_void_p(1)  # pretend a run happened; no device call is made
    with pytest.raises(L.CxxdesError):
        ens.time_unit(1, L.MILLISECONDS)
    with pytest.raises(L.CxxdesError):
        ens.time_precision(1, L.MICROSECONDS)
    ens._handle = None
