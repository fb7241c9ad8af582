"""CPU tests: the restated oracle (oracle/port) is pinned to the reference engine for M/M/1.

Two anchors: (1) committed golden vectors produced by the reference engine (tools/gen_golden.py),
which work on any machine; (2) a live comparison with oracle/_ref where the reference is present.
"""
import numpy as np
import pytest

import golden_util
import oracles
from libcxxdes_b200 import abi


def _params(case):
    return abi.MM1Params(case["lambda"], case["mu"], case["n_packets"])


@pytest.mark.parametrize("idx", range(12))
def test_port_matches_golden(port, idx):
    g = golden_util.load("mm1_reference.json")
    case = g["cases"][idx]
    got = port.run(abi.MODEL_MM1, _params(case), oracles.MM1_CLOCK, case["seed"], case["first_replication"],
                   case["n_replications"], threads=4, until=case["run_until"])
    golden_util.assert_columns_equal(got, golden_util.golden_columns(case), what=f"case {idx}")


def test_port_traces_match_golden(port):
    g = golden_util.load("mm1_reference.json")
    for tr in g["traces"]:
        p = abi.MM1Params(tr["lambda"], tr["mu"], tr["n_packets"])
        total, events = port.trace(abi.MODEL_MM1, p, oracles.MM1_CLOCK, tr["seed"], tr["replication"])
        assert total == tr["n_events"]
        assert [list(e) for e in events] == tr["trace"]


def test_event_count_identity(port):
    """n_events = 2n + W + 6: between 2n+6 and 3n+6, and exactly 4 when n = 0."""
    p = abi.MM1Params(5.0, 10.0, 400)
    c = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 1, 0, 64)
    assert (c.n_events >= 2 * 400 + 6).all() and (c.n_events <= 3 * 400 + 6).all()
    assert (c.i64[0] == 400).all()
    c0 = port.run(abi.MODEL_MM1, abi.MM1Params(5.0, 10.0, 0), oracles.MM1_CLOCK, 1, 0, 3)
    assert (c0.n_events == 4).all() and (c0.end_time == 0).all()


def test_config1_avg_latency_theory(port):
    """BASELINE config 1: one replication, 100000 packets, avg latency ~ 1/(mu - lambda) = 0.2 s."""
    c = port.run(abi.MODEL_MM1, abi.MM1Params(5.0, 10.0, 100000), oracles.MM1_CLOCK, 0, 0, 1)
    assert abs(c.f64[0][0] - 0.2) < 0.2 * 0.03
    assert 240000 < c.n_events[0] < 260000


@pytest.mark.parametrize("lam", [0.5, 5.0, 9.0, 9.9, 11.0])
def test_port_matches_reference_live(port, ref, lam):
    p = abi.MM1Params(lam, 10.0, 1500)
    a = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 0xABCDEF, 1000, 96, threads=8)
    b = ref.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 0xABCDEF, 1000, 96, threads=8)
    golden_util.assert_columns_equal(a, b, what=f"lambda={lam}")


def test_port_matches_reference_run_until(port, ref):
    p = abi.MM1Params(5.0, 10.0, 800)
    for until in (0, 1, 999, 20000, 10 ** 7):
        a = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 5, 0, 24, threads=4, until=until)
        b = ref.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 5, 0, 24, threads=4, until=until)
        golden_util.assert_columns_equal(a, b, what=f"until={until}")


def test_other_clock(port, ref):
    """unit 1 ms, precision 1 us: ticks_per_unit = 1000 again but now_seconds scales by 1e-6."""
    clk = abi.Clock.of(unit=(1, abi.MILLISECONDS), precision=(1, abi.MICROSECONDS))
    p = abi.MM1Params(5.0, 10.0, 300)
    a = port.run(abi.MODEL_MM1, p, clk, 9, 0, 16)
    b = ref.run(abi.MODEL_MM1, p, clk, 9, 0, 16)
    golden_util.assert_columns_equal(a, b, what="clock ms/us")
    clk2 = abi.Clock.of(unit=(2, abi.SECONDS), precision=(5, abi.MILLISECONDS))
    a = port.run(abi.MODEL_MM1, p, clk2, 9, 0, 16)
    b = ref.run(abi.MODEL_MM1, p, clk2, 9, 0, 16)
    golden_util.assert_columns_equal(a, b, what="clock 2s/5ms")
