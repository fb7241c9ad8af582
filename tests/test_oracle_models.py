"""CPU tests: the restated oracle is pinned to the reference engine for ALOHA, M/M/c with balking
and the memory-hierarchy model (golden vectors from the reference + live comparison where present)."""
import numpy as np
import pytest

import golden_util
import oracles
from libcxxdes_b200 import abi
from model_cases import MODELS, N_GOLDEN_CASES, clock_of


@pytest.mark.parametrize("idx", range(N_GOLDEN_CASES))
@pytest.mark.parametrize("name", sorted(MODELS))
def test_port_matches_golden(port, name, idx):
    model, cls, fname = MODELS[name]
    g = golden_util.load(fname)
    case = g["cases"][idx]
    got = port.run(model, cls(*case["params"]), clock_of(g), case["seed"], case["first_replication"],
                   case["n_replications"], threads=4, until=case["run_until"])
    golden_util.assert_columns_equal(got, golden_util.golden_columns(case), what=f"{name} case {idx}")


@pytest.mark.parametrize("name", sorted(MODELS))
def test_port_traces_match_golden(port, name):
    model, cls, fname = MODELS[name]
    g = golden_util.load(fname)
    for tr in g["traces"]:
        total, events = port.trace(model, cls(*tr["params"]), clock_of(g), tr["seed"], tr["replication"],
                                   max_events=20000)
        assert total == tr["n_events"]
        assert [list(e) for e in events] == tr["trace"]


def test_aloha_event_identity(port):
    """n_events = 2 N pps + 2 N + 3."""
    for n, pps in [(3, 2), (16, 40), (64, 10)]:
        c = port.run(abi.MODEL_ALOHA, abi.AlohaParams(n, 0, 1.0, 1.0, 1.0, 1000.0, pps), oracles.MM1_CLOCK, 1, 0, 8)
        assert (c.n_events == 2 * n * pps + 2 * n + 3).all()


def test_aloha_throughput_curve(port):
    """Pure ALOHA: S = G exp(-2G) within sampling noise (50-point sweep averaged over replications)."""
    p = abi.AlohaParams(64, 50, 0.1, 3.0, 1.0, 300.0, 0)
    c = port.run(abi.MODEL_ALOHA, p, oracles.MM1_CLOCK, 3, 0, 200, threads=8)
    g, s = c.f64[1], c.f64[0]
    for step in (4, 20, 45):
        sel = np.arange(200) % 50 == step
        G = g[sel][0]
        assert abs(s[sel].mean() - G * np.exp(-2 * G)) < 0.03


def test_mmc_conservation(port):
    p = abi.MMCParams(9.0, 1.0, 0.5, 8, 0, 500)
    c = port.run(abi.MODEL_MMC, p, oracles.MM1_CLOCK, 1, 0, 32, threads=8)
    assert (c.i64[0] + c.i64[1] == 500).all()          # served + balked = customers
    assert (c.i64[1] > 0).any()


def test_archsim_shipped_driver_counts(port):
    """basic_arch_sim as shipped: 5 events per hit, 10 per miss, + 5 for start/completion of co_main and the requester."""
    clk = abi.Clock.of((1, 0), (1, 0))
    c = port.run(abi.MODEL_ARCHSIM, abi.ArchSimParams(1, 10, 16, 32, 1, 0, 200), clk, 1, 0, 16)
    hits, misses = c.i64[0], c.i64[1]
    assert (hits + misses == 200).all()
    assert (c.n_events == 5 * hits + 10 * misses + 5).all()
    assert (c.f64[0] == hits * 1 + misses * 11).all()   # hit: 1 tick, miss: 10 + 1 ticks


@pytest.mark.parametrize("name,fields", [
    ("aloha", (24, 0, 1.5, 1.5, 1.0, 1000.0, 60)), ("aloha", (64, 50, 0.1, 3.0, 1.0, 100.0, 0)),
    ("mmc", (9.0, 1.0, 0.5, 8, 0, 400)), ("mmc", (3.0, 1.0, 0.2, 2, 0, 300)),
    ("archsim", (1, 10, 16, 32, 4, 0, 150)), ("archsim", (2, 7, 3, 16, 5, 0, 120))])
def test_port_matches_reference_live(port, ref, name, fields):
    model, cls, fname = MODELS[name]
    clk = clock_of(golden_util.load(fname))
    a = port.run(model, cls(*fields), clk, 0xFEED, 17, 48, threads=8)
    b = ref.run(model, cls(*fields), clk, 0xFEED, 17, 48, threads=8)
    golden_util.assert_columns_equal(a, b, what=f"{name} {fields}")
