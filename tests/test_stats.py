"""CPU tests of the ensemble statistics helpers (accumulator arithmetic only)."""
import numpy as np

import oracles
from libcxxdes_b200 import abi, stats
from test_sharding import _host_accumulators


def test_mean_and_ci_from_accumulators(port):
    p = abi.MM1Params(5.0, 10.0, 400)
    cols = port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 1, 0, 200, threads=4)
    acc = _host_accumulators(cols)
    mean, ci = stats.mean_and_ci(acc, 0)
    assert abs(mean - cols.f64[0].mean()) < 1e-12
    want = 1.959964 * cols.f64[0].std(ddof=1) / np.sqrt(200)
    assert abs(ci - want) < 1e-9
    assert abs(mean - 0.2) < 3 * ci + 0.02           # M/M/1: 1 / (mu - lambda)


def test_merge_equals_unsplit(port):
    p = abi.MM1Params(5.0, 10.0, 100)
    whole = _host_accumulators(port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 1, 0, 60))
    parts = [_host_accumulators(port.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 1, f, n)) for f, n in ((0, 25), (25, 35))]
    m = stats.merge(parts)
    assert m.n_replications == 60 and m.n_events == whole.n_events and m.trace_hash_sum == whole.trace_hash_sum
    assert abs(m.f64_sum[0] - whole.f64_sum[0]) < 1e-12


def test_sweep_table_shape():
    def fake(lam, mu):
        a = abi.Accumulators()
        a.n_replications = 4
        a.f64_sum[0] = 4.0 / (mu - lam)
        a.f64_sumsq[0] = 4.0 / (mu - lam) ** 2
        return a
    text = stats.mm1_sweep_table(fake, [1.0, 5.0], 10.0)
    rows = text.splitlines()
    assert rows[0].startswith("lambda, mu, rho, avg_latency") and len(rows) == 3
    assert rows[2].split(", ")[3] == rows[2].split(", ")[5]


def test_batch_means_drops_the_transient():
    rows = [{"mean": 9.0, "d_count": 5}] + [{"mean": m, "d_count": 5} for m in (1.0, 1.2, 0.8, 1.0)]
    mean, half = stats.batch_means(rows, warmup=1)
    assert abs(mean - 1.0) < 1e-12 and 0.0 < half < 0.5
