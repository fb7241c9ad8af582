"""CPU unit test of the headline kernel's per-lane state machine.

libcxxdes_b200/csrc/device/mm1_lane.cuh is a __host__ __device__ header; tests/native/mm1_lane_host.cpp
compiles it with g++ and steps single lanes through the kernel's iteration.  The result must equal the
oracle bit for bit wherever the lane did not ask for the fallback kernel (status 0)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import golden_util
import oracles
from libcxxdes_b200 import abi

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "native" / "mm1_lane_host.cpp"
OUT = ROOT / "tests" / "native" / "libmm1_lane_host.so"


@pytest.fixture(scope="module")
def lane_lib():
    hdr = ROOT / "libcxxdes_b200" / "csrc" / "device" / "mm1_lane.cuh"
    if not OUT.exists() or OUT.stat().st_mtime < max(SRC.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-mfma", "-x", "c++",
                        f"-I{ROOT / 'include'}", f"-I{ROOT / 'libcxxdes_b200' / 'csrc'}", "-shared",
                        "-o", str(OUT), str(SRC)], check=True)
    lib = C.CDLL(str(OUT))
    lib.mm1_lane_host_run.restype = C.c_int
    lib.mm1_lane_host_run.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_uint64, C.c_uint64,
                                      C.c_uint64, C.c_int64, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int64, C.c_int,
                                      C.POINTER(abi.Columns), C.POINTER(C.c_uint64)]
    return lib


def run_lanes(lib, p, n_reps, seed=0x5EED, first=0, until=-1, ring_cap=31, global_slots=0, event_slots=4,
              clock=(1000, 1, 1e-3), general_mode=0, piece=0, wide=0):
    cols = abi.HostColumns(n_reps)
    st = cols.as_struct()
    iters = C.c_uint64()
    rc = lib.mm1_lane_host_run(C.addressof(p), clock[0], clock[1], clock[2], seed, first, n_reps, until,
                               ring_cap, global_slots, event_slots, general_mode, piece, wide, C.byref(st), C.byref(iters))
    assert rc == 0
    return cols, iters.value


def masked_equal(got, want, what):
    ok = (got.status & ~np.uint32(abi.REP_UNFINISHED)) == 0
    for name in ("end_time", "n_events", "trace_hash"):
        assert np.array_equal(getattr(got, name)[ok], getattr(want, name)[ok]), f"{what}: {name}"
    for k in range(2):
        assert np.array_equal(got.f64[k][ok].view(np.uint64), want.f64[k][ok].view(np.uint64)), f"{what}: f64[{k}]"
        assert np.array_equal(got.i64[k][ok], want.i64[k][ok]), f"{what}: i64[{k}]"
    return ok


@pytest.mark.parametrize("lam,n_packets,n_reps", [(5.0, 2000, 300), (0.5, 1000, 200), (9.0, 1500, 200),
                                                  (12.0, 300, 100), (5.0, 2, 50), (5.0, 3, 50), (5.0, 7, 50)])
def test_lane_matches_oracle(lane_lib, port, lam, n_packets, n_reps):
    p = abi.MM1Params(lam, 10.0, n_packets)
    got, _ = run_lanes(lane_lib, p, n_reps)
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 0x5EED, 0, n_reps, threads=oracles.os.cpu_count())
    ok = masked_equal(got, want, f"lambda={lam} n={n_packets}")
    if lam <= 5.0:
        assert ok.all()                       # the 32-entry ring holds every queue at rho <= 0.5 here
    bad = got.status[~ok]
    assert ((bad & (abi.REP_QUEUE_OVERFLOW | abi.REP_TIME_RANGE)) != 0).all()   # only "redo me" flags


@pytest.mark.parametrize("event_slots,ring_cap", [(1, 31), (2, 8), (3, 5), (6, 31), (4, 2), (4, 3)])
def test_slot_and_ring_shapes_do_not_change_results(lane_lib, port, event_slots, ring_cap):
    p = abi.MM1Params(6.0, 10.0, 600)
    got, _ = run_lanes(lane_lib, p, 150, seed=3, ring_cap=ring_cap, event_slots=event_slots)
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 3, 0, 150, threads=4)
    ok = masked_equal(got, want, f"slots={event_slots} cap={ring_cap}")
    if ring_cap >= 31:
        assert ok.mean() > 0.9
    # a replication is flagged exactly when its queue outgrew the ring (the pair-wise generation
    # may give up one entry early)
    assert (want.i64[1][~ok] >= ring_cap - 1).all()
    assert (want.i64[1][ok] <= ring_cap).all()


@pytest.mark.parametrize("until", [0, 1, 999, 20000, 10 ** 7])
def test_lane_run_until(lane_lib, port, until):
    p = abi.MM1Params(5.0, 10.0, 800)
    got, _ = run_lanes(lane_lib, p, 128, seed=5, until=until)
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 5, 0, 128, threads=4, until=until)
    assert masked_equal(got, want, f"until={until}").all()


def test_lane_tick_range_guard(lane_lib, port):
    """Slow arrivals (every delay still below 2^28 ticks, the launcher's precondition): the clock
    passes 2^30 and the lane must ask for the 64-bit kernel instead of wrapping."""
    p = abi.MM1Params(0.0002, 0.01, 400)
    got, _ = run_lanes(lane_lib, p, 64)
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 0x5EED, 0, 64, threads=4)
    ok = masked_equal(got, want, "slow arrivals")
    assert (got.status[~ok] & abi.REP_TIME_RANGE).all()
    assert (want.end_time[ok] < 2 ** 30).all()
    assert (~ok).sum() > 32


def test_lane_other_clocks(lane_lib, port):
    p = abi.MM1Params(5.0, 10.0, 300)
    for unit, prec, folded in [((1, abi.MILLISECONDS), (1, abi.MICROSECONDS), (1000, 1, 1e-6)),
                               ((2, abi.SECONDS), (5, abi.MILLISECONDS), (400, 5, 1e-3))]:
        got, _ = run_lanes(lane_lib, p, 64, seed=9, clock=folded)
        want = port.run(abi.MODEL_MM1, p, abi.Clock.of(unit, prec), 9, 0, 64)
        assert masked_equal(got, want, f"clock {unit}/{prec}").all()


@pytest.mark.parametrize("lam,n_packets,slots", [(9.0, 3000, 1024), (9.9, 2000, 2048), (12.0, 400, 512), (5.0, 800, 64),
                                                   (9.5, 1500, 64)])
def test_spilling_ring_matches_oracle(lane_lib, port, lam, n_packets, slots):
    """Heavily loaded queues: entries leave the 31-packet window for the second ring level and come
    back when the consumer reaches them; nothing but a queue beyond that level may be flagged."""
    p = abi.MM1Params(lam, 10.0, n_packets)
    got, _ = run_lanes(lane_lib, p, 120, seed=21, global_slots=slots)
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 21, 0, 120, threads=oracles.os.cpu_count())
    ok = masked_equal(got, want, f"spill lambda={lam}")
    assert (want.i64[1] > 40).mean() > 0.3 or lam <= 5.0     # the scenario really leaves the window
    assert (want.i64[1][~ok] >= slots - 2).all()             # flagged only when the second level is full too
    assert (got.status[~ok] & abi.REP_QUEUE_OVERFLOW).all()
    if slots >= 1024:
        assert ok.all()


@pytest.mark.parametrize("lam,n_packets,piece,until,slots", [
    (5.0, 800, 7000, -1, 0), (5.0, 800, 100, 30000, 0), (5.0, 800, 1, 300, 0), (0.5, 300, 2500, -1, 0),
    (9.0, 1500, 9000, -1, 1024), (9.5, 1000, 333, 90000, 256), (5.0, 9, 50, -1, 0), (5.0, 2, 1, -1, 0)])
def test_lane_run_in_pieces(lane_lib, port, lam, n_packets, piece, until, slots):
    """run_for(piece) repeatedly == one run (environment.ipp:190-214): the lane is saved at every
    horizon, restored into an empty ring and its ring entries are drawn again."""
    p = abi.MM1Params(lam, 10.0, n_packets)
    got, _ = run_lanes(lane_lib, p, 96, seed=77, until=until, piece=piece, global_slots=slots)
    want = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 77, 0, 96, threads=4, until=until)
    assert masked_equal(got, want, f"pieces of {piece}").all()


US = (1000000, 1, 1e-6)     # time_unit 1 s, time_precision 1 us, folded as ensemble.py does


@pytest.mark.parametrize("lam,n_packets,until,piece,slots", [
    (5.0, 12000, -1, 0, 0),               # clock ends at 2.4e9 ticks: two moves of the epoch
    (5.0, 12000, 1500000000, 0, 0),       # horizon beyond 2^30
    (5.0, 9000, -1, 400000000, 0),        # run_for pieces across the moves of the epoch
    (9.0, 14000, -1, 0, 1024),            # spilling ring: entries in the second level are rebased too
    (9.0, 14000, 1300000000, 350000000, 1024),
    (5.0, 300, -1, 0, 0)])                # short run: wide mode without ever moving the epoch
def test_lane_wide_clock_matches_oracle(lane_lib, port, lam, n_packets, until, piece, slots):
    """1 us precision with a 1 s unit: the clock passes 2^30 ticks; the lane keeps its 32-bit times relative
    to an epoch that it moves forward (MODE_WIDE) and must still match the oracle bit for bit."""
    p = abi.MM1Params(lam, 10.0, n_packets)
    got, _ = run_lanes(lane_lib, p, 48, seed=19, until=until, piece=piece, global_slots=slots, clock=US, wide=1)
    want = port.run(abi.MODEL_MM1, p, abi.Clock.of((1, abi.SECONDS), (1, abi.MICROSECONDS)), 19, 0, 48,
                    threads=oracles.os.cpu_count(), until=until)
    assert masked_equal(got, want, "wide clock").all()
    if until < 0 and n_packets > 5000:
        assert want.end_time.max() > 2 ** 30


def test_wide_mode_equals_plain_mode_on_a_small_clock(lane_lib):
    p = abi.MM1Params(5.0, 10.0, 700)
    a, _ = run_lanes(lane_lib, p, 64, seed=3)
    b, _ = run_lanes(lane_lib, p, 64, seed=3, wide=1)
    golden_util.assert_columns_equal(a, b, what="wide vs plain")


def test_compile_time_modes_agree(lane_lib):
    """The specialised variants (no horizon, tick_mult == 1) compute what the general one does."""
    p = abi.MM1Params(5.0, 10.0, 500)
    for until, clock in [(-1, (1000, 1, 1e-3)), (30000, (1000, 1, 1e-3)), (-1, (400, 5, 1e-3)), (9000, (400, 5, 1e-3))]:
        a, _ = run_lanes(lane_lib, p, 64, seed=11, until=until, clock=clock)
        b, _ = run_lanes(lane_lib, p, 64, seed=11, until=until, clock=clock, general_mode=1)
        golden_util.assert_columns_equal(a, b, what=f"until={until} clock={clock}")


def test_lane_efficiency_at_the_headline_load(lane_lib):
    """Two packets per iteration is the design point: at rho = 0.5 a lane must need barely more than
    n_packets / 2 iterations (start-up and wind-down included)."""
    p = abi.MM1Params(5.0, 10.0, 10000)
    _, iters = run_lanes(lane_lib, p, 20)
    assert iters / 20 < 10000 / 2 * 1.02 + 20
