"""cxxdes_b200_trace on the GPU: the event trace of one replication -- (time, priority, kind, process) of every token
environment::step pops (environment.ipp:117-146) -- against the full traces the unmodified reference engine produced
(tests/golden/*_reference.json "traces", tools/gen_golden.py) and against the CPU oracle."""
import os

import numpy as np
import pytest

import golden_util
import kat_programs as K
import oracles
import libcxxdes_b200 as L
from libcxxdes_b200 import abi
from model_cases import MODELS, clock_of, clock_pairs

pytestmark = pytest.mark.gpu


def _ensemble(params, n, seed, clock, first=0):
    ens = L.Ensemble(params, n, seed=seed, first_replication=first)
    ens.time_unit(*clock[0]).time_precision(*clock[1])
    return ens


def test_mm1_traces_match_the_reference_engine(built):
    g = golden_util.load("mm1_reference.json")
    assert g["traces"]
    for tr in g["traces"]:
        p = abi.MM1Params(tr["lambda"], tr["mu"], tr["n_packets"])
        ens = _ensemble(p, tr["replication"] + 3, tr["seed"], ((1, L.SECONDS), (1, L.MILLISECONDS)))
        total, digest, events = ens.trace(tr["replication"], max_events=tr["n_events"] + 8)
        cols = ens.run().fetch()
        ens.close()
        assert total == tr["n_events"] == int(cols.n_events[tr["replication"]])
        assert [list(e) for e in events] == tr["trace"]
        # the digest of the traced run is the digest the ensemble's (fused, 32-bit) kernel produced for this replication
        assert digest == int(cols.trace_hash[tr["replication"]])


@pytest.mark.parametrize("name", sorted(MODELS))
def test_model_traces_match_the_reference_engine(built, name):
    model, cls, fname = MODELS[name]
    g = golden_util.load(fname)
    assert g["traces"]
    for tr in g["traces"]:
        ens = _ensemble(cls(*tr["params"]), tr["replication"] + 2, tr["seed"], clock_pairs(g))
        total, digest, events = ens.trace(tr["replication"], max_events=20000)
        cols = ens.run().fetch()
        ens.close()
        assert total == tr["n_events"]
        assert [list(e) for e in events] == tr["trace"]
        assert digest == int(cols.trace_hash[tr["replication"]])


def test_trace_is_truncated_and_follows_horizons_and_shards(port):
    """max_events smaller than the run, `until` as run_until, and a shard that does not start at replication 0."""
    p = abi.MM1Params(5.0, 10.0, 200)
    clock = ((1, L.SECONDS), (1, L.MILLISECONDS))
    ens = _ensemble(p, 16, 11, clock, first=1000)
    want_total, want = port.trace(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 11, 1005, max_events=4096)
    total, _, events = ens.trace(5, max_events=4096)
    assert total == want_total and events == want
    total, _, events = ens.trace(5, max_events=10)
    assert total == want_total and events == want[:10]
    want_total, want = port.trace(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 11, 1005, max_events=4096, until=7000)
    total, digest, events = ens.trace(5, max_events=4096, until=7000)
    assert total == want_total and events == want and all(e[0] <= 7000 for e in events)
    part = ens.run_until(7000).fetch(strict=False)
    assert digest == int(part.trace_hash[5]) and total == int(part.n_events[5])
    with pytest.raises(L.CxxdesError):
        ens.trace(16)
    ens.close()


@pytest.mark.parametrize("k", sorted(K.SCENARIOS))
def test_program_traces_match_the_reference_engine(built, k):
    """The eighteen scenarios of the reference's own tests and examples as process programs: the trace of the full-size
    interpreter equals, row for row, the trace the reference engine produced for the scenario written natively on it
    (tests/golden/kat_reference.json); processes are named by their ordinals."""
    g = golden_util.load("kat_reference.json")
    case = next(c for c in g["cases"] if c["scenario"] == k)
    prog, until = K.SCENARIOS[k]()
    ens = _ensemble(prog, 1, 1, ((1, L.SECONDS), (1, L.SECONDS)))
    total, digest, events = ens.trace(0, max_events=1000, until=until)
    cols = (ens.run_until(until) if until >= 0 else ens.run()).fetch(strict=False)
    ens.close()
    assert total == case["n_events"] and [list(e) for e in events] == case["trace"]
    assert digest == int(cols.trace_hash[0])
