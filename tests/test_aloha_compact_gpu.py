"""ALOHA on 4-byte tokens (aloha_compact_kernel, device/engine32.cuh heap24x: time relative to a moving epoch + 8-bit code,
16-bit station counters; the first pass wherever it applies) and on 8-byte tokens (aloha_fast_kernel,
CXXDES_B200_FLAG_WIDE_TOKENS): both against the reference's golden vectors and the oracle, bit for bit, against each other
at a size that makes every lane steal work and rebase its epoch many times, and the hand-over of the replications whose
waiting time does not fit 24 bits (REP_TIME_RANGE -> 64-bit kernel).  The heap's moves are checked against
std::priority_queue on the CPU by tests/native/device_heap_host.cpp (test_shared_math.py)."""
import os

import numpy as np
import pytest

import golden_util
import libcxxdes_b200 as L
from libcxxdes_b200 import abi
from model_cases import MODELS, N_GOLDEN_CASES, clock_of, clock_pairs

pytestmark = pytest.mark.gpu

S_MS = ((1, L.SECONDS), (1, L.MILLISECONDS))


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
    passes = ens.pass_counts()
    ens.close()
    return cols, acc, name, passes


@pytest.mark.parametrize("flags,kernel", [(0, "aloha_compact_kernel<false>"), (abi.FLAG_WIDE_TOKENS, "aloha_fast_kernel<false>")])
@pytest.mark.parametrize("idx", range(N_GOLDEN_CASES))
def test_both_token_widths_match_reference_golden(built, idx, flags, kernel):
    model, cls, fname = MODELS["aloha"]
    g = golden_util.load(fname)
    case = g["cases"][idx]
    got, _, name, _ = _run(cls(*case["params"]), case["n_replications"], clock_pairs(g), case["seed"],
                           first=case["first_replication"], until=case["run_until"], flags=flags)
    assert name == kernel
    assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
    golden_util.assert_columns_equal(got, golden_util.golden_columns(case), what=f"aloha golden {idx} on {kernel}")


@pytest.mark.parametrize("fields,n_reps", [
    ((64, 50, 0.1, 3.0, 1.0, 100.0, 0), 2000),
    ((64, 0, 3.0, 3.0, 1.0, 1000.0, 40), 512)])
def test_wide_tokens_match_oracle(port, fields, n_reps):
    """(the default path -- 4-byte tokens -- has these cases in test_models_gpu.py::test_gpu_matches_oracle)"""
    model, cls, fname = MODELS["aloha"]
    g = golden_util.load(fname)
    p = cls(*fields)
    got, acc, name, _ = _run(p, n_reps, clock_pairs(g), 0xFEED, first=17, flags=abi.FLAG_WIDE_TOKENS)
    want = port.run(model, p, clock_of(g), 0xFEED, 17, n_reps, threads=os.cpu_count())
    assert name == "aloha_fast_kernel<false>" and not got.status.any()
    golden_util.assert_columns_equal(got, want, what=f"aloha {fields} on 8-byte tokens")
    assert acc.trace_hash_sum == int(want.trace_hash.sum(dtype=np.uint64))


def test_token_widths_agree_on_a_large_sweep(built):
    """BASELINE config 3's sweep (G = 0.1 .. 3.0, 64 stations, G * 1000 packets per station), 80 000 replications: more than
    the 75 776 resident lanes, so lanes claim second replications; the low-load replications run to 6.4e7 ticks = 15
    rebases of the 2^22-tick epoch each.  Same columns from both kernels."""
    p = abi.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0)
    a, acc_a, name_a, _ = _run(p, 80000, S_MS, 0x5EED, first=1 << 18)
    b, acc_b, name_b, _ = _run(p, 80000, S_MS, 0x5EED, first=1 << 18, flags=abi.FLAG_WIDE_TOKENS)
    assert (name_a, name_b) == ("aloha_compact_kernel<false>", "aloha_fast_kernel<false>")
    assert not a.status.any() and not b.status.any()
    golden_util.assert_columns_equal(a, b, what="aloha 4-byte against 8-byte tokens")
    assert acc_a.trace_hash_sum == acc_b.trace_hash_sum and acc_a.n_events == acc_b.n_events
    assert int(a.end_time.max()) > 15 << 22


def test_waiting_times_beyond_24_bits_are_handed_over(port, monkeypatch):
    """With the selection threshold lowered (CXXDES_B200_ALOHA_TAIL = 3: the kernel is chosen although a waiting time
    passes heap24x::DELAY_LIMIT with probability ~e^-3.6 per draw), most replications raise REP_TIME_RANGE on the first
    pass and are redone by the 64-bit kernel; the columns are the oracle's all the same, nothing is flagged."""
    model, cls, fname = MODELS["aloha"]
    g = golden_util.load(fname)
    monkeypatch.setenv("CXXDES_B200_ALOHA_TAIL", "3")
    # mean wait = N / lambda = 3.5e6 ticks of 1 ms; DELAY_LIMIT = 1.26e7 ticks
    p = cls(7, 0, 0.002, 0.002, 1.0, 1000.0, 6)
    got, acc, name, passes = _run(p, 3000, clock_pairs(g), 31, first=5)
    want = port.run(model, p, clock_of(g), 31, 5, 3000, threads=os.cpu_count())
    assert name == "aloha_compact_kernel<false>" and not got.status.any()
    golden_util.assert_columns_equal(got, want, what="aloha with hand-over to the 64-bit kernel")
    assert 1000 < passes["fallback"] < 2900          # waits that long did occur, and many replications never drew one


def test_pieces_on_4_byte_tokens_with_hand_over(port, monkeypatch):
    """run_for in pieces where some replications are saved on 4-byte tokens (epoch in the saved words), some finished,
    some handed to the 64-bit kernel (phase REPLAY): after every piece the oracle's run_until columns."""
    model, cls, fname = MODELS["aloha"]
    g = golden_util.load(fname)
    monkeypatch.setenv("CXXDES_B200_ALOHA_TAIL", "8")
    p = cls(12, 0, 0.01, 0.01, 1.0, 1000.0, 25)    # mean wait 1.2e6 ticks: e^-10.5 per draw, 600 draws per replication
    n_reps = 4000
    clock = clock_pairs(g)
    ends = port.run(model, p, clock_of(g), 77, 0, n_reps, threads=os.cpu_count()).end_time
    piece = int(ends.max()) // 7
    ens = L.Ensemble(p, n_reps, seed=77)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    assert ens.kernel_name() == "aloha_compact_kernel<false>"
    for k in range(1, 9):
        ens.run_for(piece)
        got = ens.fetch(strict=False)
        want = port.run(model, p, clock_of(g), 77, 0, n_reps, threads=os.cpu_count(), until=k * piece)
        assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
        golden_util.assert_columns_equal(got, want, what=f"aloha (4-byte tokens) after piece {k}")
    assert not ens.fetch(strict=False).status.any()
    assert 5 <= ens.pass_counts()["fallback"] <= 200          # the replications that are replayed on every piece
    ens.close()
