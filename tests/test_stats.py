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


def test_aloha_sweep_table_and_writers(port, tmp_path):
    """examples/aloha.cpp:93-109 from one ensemble: rows follow the sweep, S tracks G exp(-2G) for pure ALOHA."""
    import json
    import golden_util
    from model_cases import MODELS, clock_of
    model, cls, fname = MODELS["aloha"]
    g = golden_util.load(fname)
    steps = 5
    p = cls(16, steps, 0.2, 1.0, 1.0, 200.0, 0)
    cols = port.run(model, p, clock_of(g), 3, 0, 40, threads=4)
    rows = stats.aloha_sweep_table(cols, steps).splitlines()
    assert rows[0].startswith("G [Erlang], S [Erlang], ci95") and len(rows) == 1 + steps
    prev = 0.0
    for line in rows[1:]:
        G, S, half, lo, hi, n = (float(x) for x in line.split(", "))
        assert G > prev and n == 8 and lo <= S <= hi
        assert abs(S - G * np.exp(-2 * G)) < 0.08
        prev = G
    acc = _host_accumulators(cols)
    js = stats.summary(acc, names_f64=("S", "G"), names_i64=("successful", "packets_per_station"), run_ms=2.0)
    assert js["n_replications"] == 40 and js["events_per_s"] == acc.n_events / 2e-3
    assert abs(js["S"]["mean"] - cols.f64[0].mean()) < 1e-12 and js["successful"]["sum"] == int(cols.i64[0].sum())
    stats.write_json(tmp_path / "run.json", js)
    assert json.loads((tmp_path / "run.json").read_text())["trace_hash_sum"] == js["trace_hash_sum"]
    stats.write_columns_csv(tmp_path / "cols.csv", cols, names_f64=("S", "G"))
    lines = (tmp_path / "cols.csv").read_text().splitlines()
    assert len(lines) == 41 and lines[0].split(", ")[5] == "S"
    r7 = lines[8].split(", ")
    assert int(r7[1]) == cols.end_time[7] and float(r7[5]) == cols.f64[0][7] and int(r7[3], 16) == cols.trace_hash[7]


def test_sweep_table_groups_by_point(port):
    import kat_programs as K
    lams = [2.0, 5.0, 8.0]
    prog = K.mm1_program(lams, 10.0, 400)
    cols = port.run(abi.MODEL_PROGRAM, prog.to_params(), oracles.MM1_CLOCK, 5, 1, 60, threads=4)
    rows = stats.sweep_table(cols, 3, stat=1, scale=1.0 / 400, labels=lams, first_replication=1, name="latency").splitlines()
    assert rows[0] == "point, latency, ci95, replications" and len(rows) == 4
    means = [float(r.split(", ")[1]) for r in rows[1:]]
    assert [r.split(", ")[0] for r in rows[1:]] == ["2.0", "5.0", "8.0"] and means[0] < means[1] < means[2]
    assert abs(means[1] - 0.2) < 0.03 and all(r.endswith(", 20") for r in rows[1:])
