"""ALOHA and M/M/c on 4-byte tokens (device/engine32.cuh heap_rel: time relative to a moving epoch + a short code; ALOHA
24 + 8 bits with 16-bit station counters, M/M/c 20 + 12 bits; the first pass wherever it applies) and on 8-byte tokens
(CXXDES_B200_FLAG_WIDE_TOKENS): both against the reference's golden vectors and the oracle, bit for bit, against each
other at sizes that make every lane steal work and rebase its epoch, run_for in pieces (saved state holds absolute times),
and the hand-over of the replications whose delays do not fit the relative time (REP_TIME_RANGE -> 64-bit kernel).  The
heap's moves are checked against std::priority_queue on the CPU by tests/native/device_heap_host.cpp (test_shared_math.py)."""
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


# ---- M/M/c ----------------------------------------------------------------------------------------------------------

MMC_KERNELS = [(0, "mmc_fast_kernel<false,narrow>"), (abi.FLAG_WIDE_TOKENS, "mmc_fast_kernel<false,wide>")]


@pytest.mark.parametrize("flags,kernel", MMC_KERNELS)
@pytest.mark.parametrize("idx", range(N_GOLDEN_CASES))
def test_mmc_both_token_widths_match_reference_golden(built, idx, flags, kernel):
    model, cls, fname = MODELS["mmc"]
    g = golden_util.load(fname)
    case = g["cases"][idx]
    got, _, name, _ = _run(cls(*case["params"]), case["n_replications"], clock_pairs(g), case["seed"],
                           first=case["first_replication"], until=case["run_until"], flags=flags)
    assert name == kernel
    assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
    golden_util.assert_columns_equal(got, golden_util.golden_columns(case), what=f"mmc golden {idx} on {kernel}")


@pytest.mark.parametrize("fields,n_reps", [((9.0, 1.0, 0.5, 8, 0, 400), 1024), ((12.0, 1.0, 0.5, 8, 0, 300), 1024),
                                           ((3.0, 1.0, 0.2, 2, 0, 300), 512)])
def test_mmc_wide_tokens_match_oracle(port, fields, n_reps):
    """(the default path -- 4-byte tokens -- has these cases in test_models_gpu.py::test_gpu_matches_oracle)"""
    model, cls, fname = MODELS["mmc"]
    g = golden_util.load(fname)
    p = cls(*fields)
    got, acc, name, _ = _run(p, n_reps, clock_pairs(g), 0xFEED, first=17, flags=abi.FLAG_WIDE_TOKENS)
    want = port.run(model, p, clock_of(g), 0xFEED, 17, n_reps, threads=os.cpu_count())
    assert name == "mmc_fast_kernel<false,wide>" and not got.status.any()
    golden_util.assert_columns_equal(got, want, what=f"mmc {fields} on 8-byte tokens")
    assert acc.trace_hash_sum == int(want.trace_hash.sum(dtype=np.uint64))


def test_mmc_token_widths_agree_at_full_length(built):
    """BASELINE config 4's parameters (1000 customers: ordinals up to 2001 of the 2046 that fit), 200 000 replications =
    more than the 85 248 resident lanes of the narrow-token kernel.  Same columns from both kernels."""
    p = abi.MMCParams(9.0, 1.0, 0.5, 8, 0, 1000)
    a, acc_a, name_a, pa = _run(p, 200000, S_MS, 0x5EED, first=1 << 19)
    b, acc_b, name_b, _ = _run(p, 200000, S_MS, 0x5EED, first=1 << 19, flags=abi.FLAG_WIDE_TOKENS)
    assert (name_a, name_b) == ("mmc_fast_kernel<false,narrow>", "mmc_fast_kernel<false,wide>")
    assert not a.status.any() and not b.status.any()
    golden_util.assert_columns_equal(a, b, what="mmc 4-byte against 8-byte tokens")
    assert acc_a.trace_hash_sum == acc_b.trace_hash_sum and acc_a.n_events == acc_b.n_events
    assert pa["fallback"] < 20


def test_mmc_more_customers_than_the_narrow_ordinals(port):
    """1023 customers need ordinal 2047 = the narrow tokens' no-target code: the 8-byte kernel runs the model."""
    model, cls, fname = MODELS["mmc"]
    g = golden_util.load(fname)
    p = cls(9.0, 1.0, 0.5, 8, 0, 1023)
    got, _, name, _ = _run(p, 64, clock_pairs(g), 3)
    want = port.run(model, p, clock_of(g), 3, 0, 64, threads=os.cpu_count())
    assert name == "mmc_fast_kernel<false,wide>" and not got.status.any()
    golden_util.assert_columns_equal(got, want, what="mmc with 1023 customers")
    p = cls(9.0, 1.0, 0.5, 8, 0, 1022)
    got, _, name, _ = _run(p, 64, clock_pairs(g), 3)
    want = port.run(model, p, clock_of(g), 3, 0, 64, threads=os.cpu_count())
    assert name == "mmc_fast_kernel<false,narrow>" and not got.status.any()
    golden_util.assert_columns_equal(got, want, what="mmc with 1022 customers")


def test_mmc_delays_beyond_20_bits_are_handed_over(port, monkeypatch):
    """Selection threshold lowered (CXXDES_B200_MMC_TAIL = 3): service and inter-arrival times pass heap_rel<.., 12>::
    DELAY_LIMIT = 786 432 ticks with probability 2-3 % per draw; most replications are redone by the 64-bit kernel, the
    clock passes 2^20 ticks (rebases) in all of them; columns are the oracle's."""
    model, cls, fname = MODELS["mmc"]
    g = golden_util.load(fname)
    monkeypatch.setenv("CXXDES_B200_MMC_TAIL", "3")
    p = cls(0.005, 0.0045, 100.0, 2, 0, 40)
    got, acc, name, passes = _run(p, 3000, clock_pairs(g), 31, first=5)
    want = port.run(model, p, clock_of(g), 31, 5, 3000, threads=os.cpu_count())
    assert name == "mmc_fast_kernel<false,narrow>" and not got.status.any()
    golden_util.assert_columns_equal(got, want, what="mmc with hand-over to the 64-bit kernel")
    assert 1000 < passes["fallback"] < 2900
    assert int(want.end_time.min()) > 1 << 20


def test_mmc_pieces_on_4_byte_tokens_with_hand_over(port, monkeypatch):
    """run_for in pieces: replications saved on narrow tokens (absolute times in the saved words, the epoch restarts at
    the restored clock), finished ones, and ones handed to the 64-bit kernel; after every piece the oracle's columns."""
    model, cls, fname = MODELS["mmc"]
    g = golden_util.load(fname)
    monkeypatch.setenv("CXXDES_B200_MMC_TAIL", "8")
    p = cls(0.012, 0.011, 60.0, 2, 0, 60)      # mean service 90 909 ticks: e^-8.65 per draw
    n_reps = 4000
    clock = clock_pairs(g)
    ends = port.run(model, p, clock_of(g), 77, 0, n_reps, threads=os.cpu_count()).end_time
    piece = int(ends.max()) // 7
    ens = L.Ensemble(p, n_reps, seed=77)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    assert ens.kernel_name() == "mmc_fast_kernel<false,narrow>"
    for k in range(1, 9):
        ens.run_for(piece)
        got = ens.fetch(strict=False)
        want = port.run(model, p, clock_of(g), 77, 0, n_reps, threads=os.cpu_count(), until=k * piece)
        assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
        golden_util.assert_columns_equal(got, want, what=f"mmc (4-byte tokens) after piece {k}")
    assert not ens.fetch(strict=False).status.any()
    assert 5 <= ens.pass_counts()["fallback"] <= 400
    ens.close()


# ---- memory hierarchy: 32-bit stamps / 16-bit addresses / bare wait lists in shared memory ---------------------------

@pytest.mark.parametrize("flags", [0, abi.FLAG_WIDE_TOKENS])
@pytest.mark.parametrize("idx", range(N_GOLDEN_CASES))
def test_archsim_both_layouts_match_reference_golden(built, idx, flags):
    model, cls, fname = MODELS["archsim"]
    g = golden_util.load(fname)
    case = g["cases"][idx]
    got, _, name, _ = _run(cls(*case["params"]), case["n_replications"], clock_pairs(g), case["seed"],
                           first=case["first_replication"], until=case["run_until"], flags=flags)
    assert name.endswith(",narrow>") == (flags == 0), name
    assert not (got.status & ~np.uint32(abi.REP_UNFINISHED)).any()
    golden_util.assert_columns_equal(got, golden_util.golden_columns(case), what=f"archsim golden {idx} on {name}")


@pytest.mark.parametrize("fields,n_reps", [((1, 10, 16, 32, 1, 0, 1024), 1024), ((1, 10, 16, 32, 4, 0, 256), 1024),
                                           ((0, 0, 4, 8, 4, 0, 100), 512), ((2, 7, 3, 16, 5, 0, 120), 512)])
def test_archsim_wide_layout_matches_oracle(port, fields, n_reps):
    """(the default -- narrow -- layout has these cases in test_models_gpu.py::test_gpu_matches_oracle)"""
    model, cls, fname = MODELS["archsim"]
    g = golden_util.load(fname)
    p = cls(*fields)
    got, acc, name, _ = _run(p, n_reps, clock_pairs(g), 0xFEED, first=17, flags=abi.FLAG_WIDE_TOKENS)
    want = port.run(model, p, clock_of(g), 0xFEED, 17, n_reps, threads=os.cpu_count())
    assert "narrow" not in name and not got.status.any()
    golden_util.assert_columns_equal(got, want, what=f"archsim {fields} on the wide layout")


def test_archsim_layouts_agree_at_size_and_in_pieces(port):
    """4 requesters x 256 loads, 150 000 replications (more than the 113 664 resident lanes of the narrow layout): same
    columns from both layouts; then run_for in pieces on the wide layout against the oracle (the narrow one has
    test_models_gpu.py::test_archsim_run_for_in_many_pieces...)."""
    p = abi.ArchSimParams(1, 10, 16, 32, 4, 0, 256)
    S_S = ((1, L.SECONDS), (1, L.SECONDS))
    a, acc_a, name_a, _ = _run(p, 150000, S_S, 0x5EED, first=1 << 21)
    b, acc_b, name_b, _ = _run(p, 150000, S_S, 0x5EED, first=1 << 21, flags=abi.FLAG_WIDE_TOKENS)
    assert name_a == "archsim_kernel<k requesters,narrow>" and name_b == "archsim_kernel<k requesters>"
    assert not a.status.any() and not b.status.any()
    golden_util.assert_columns_equal(a, b, what="archsim narrow against wide layout")
    assert acc_a.trace_hash_sum == acc_b.trace_hash_sum
    model, cls, fname = MODELS["archsim"]
    g = golden_util.load(fname)
    q = cls(1, 10, 16, 32, 4, 0, 96)
    clock = clock_pairs(g)
    ends = port.run(model, q, clock_of(g), 4, 0, 300, threads=os.cpu_count()).end_time
    piece = max(1, int(ends.max()) // 5)
    ens = L.Ensemble(q, 300, seed=4, flags=abi.FLAG_WIDE_TOKENS)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    for k in range(1, 7):
        ens.run_for(piece)
        want = port.run(model, q, clock_of(g), 4, 0, 300, threads=os.cpu_count(), until=k * piece)
        golden_util.assert_columns_equal(ens.fetch(strict=False), want, what=f"archsim (wide layout) after piece {k}")
    ens.close()


def test_archsim_long_runs_keep_64_bit_stamps(built):
    """n_requesters * n_loads * (l1 + l2 + 1) >= 2^31 ticks could pass 32 bits: the wide layout runs the model."""
    p = abi.ArchSimParams(1000000, 1000000, 16, 32, 2, 0, 800)
    ens = L.Ensemble(p, 64, seed=1)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.SECONDS)
    ens.run()
    got = ens.fetch()
    assert ens.kernel_name() == "archsim_kernel<k requesters>"
    assert not got.status.any() and int(got.end_time.max()) > 1 << 31
    ens.close()
