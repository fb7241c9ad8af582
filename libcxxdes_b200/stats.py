"""Ensemble statistics: what the `main()` of the reference examples does after `sim.run()`
(examples/producer_consumer.cpp:67-73 prints one avg_latency per load point; aloha.cpp:93-109 one
(G, S) pair), extended to an ensemble: means with confidence intervals from the on-device
accumulators, quantiles from the fetched columns."""
import math

import numpy as np

from . import abi


def mean_and_ci(acc, index, z=1.959964):
    """Mean of f64[index] over the ensemble and the half-width of its normal-approximation CI,
    from the reduced accumulators alone (no per-replication data leaves the GPU)."""
    n = acc.n_replications
    if n == 0:
        return float("nan"), float("nan")
    mean = acc.f64_sum[index] / n
    if n < 2:
        return mean, float("nan")
    var = max(0.0, (acc.f64_sumsq[index] - n * mean * mean) / (n - 1))
    return mean, z * math.sqrt(var / n)


def merge(accs):
    """Sum accumulator structs of several shards (what the NCCL all-reduce does on the device)."""
    out = abi.Accumulators()
    for a in accs:
        out.n_replications += a.n_replications
        out.n_events += a.n_events
        out.n_errors += a.n_errors
        out.trace_hash_sum = (out.trace_hash_sum + a.trace_hash_sum) & (2 ** 64 - 1)
        out.end_time_sum += a.end_time_sum
        for k in range(abi.N_I64):
            out.i64_sum[k] += a.i64_sum[k]
        for k in range(abi.N_F64):
            out.f64_sum[k] += a.f64_sum[k]
            out.f64_sumsq[k] += a.f64_sumsq[k]
    return out


def quantiles(column, qs=(0.05, 0.5, 0.95)):
    return [float(x) for x in np.quantile(np.asarray(column, dtype=np.float64), qs)]


def mm1_sweep_table(run_point, lambdas, mu):
    """The CSV of examples/producer_consumer.cpp:67-73 with an ensemble behind every row.
    `run_point(lambda, mu)` returns the accumulators of one load point."""
    lines = ["lambda, mu, rho, avg_latency, ci95, theory"]
    for lam in lambdas:
        acc = run_point(lam, mu)
        mean, ci = mean_and_ci(acc, 0)
        lines.append(f"{lam:.3f}, {mu:.3f}, {lam / mu:.3f}, {mean:.6f}, {ci:.6f}, {1.0 / (mu - lam):.6f}")
    return "\n".join(lines)


def aloha_sweep_table(columns, sweep_steps, qs=(0.05, 0.95)):
    """The table of examples/aloha.cpp:93-109 ("G [Erlang], S [Erlang]", one row per offered load) from ONE ensemble
    run: replication r ran load point r mod sweep_steps (cxxdes_b200_aloha_params.sweep_steps), f64[0] = S and
    f64[1] = G as the example computes them.  Every row is the mean over the replications of its load point, with
    the half-width of the 95 % interval and two quantiles of S next to it."""
    S, G = np.asarray(columns.f64[0]), np.asarray(columns.f64[1])
    lines = ["G [Erlang], S [Erlang], ci95, " + ", ".join(f"q{int(round(100 * q)):02d}" for q in qs) + ", replications"]
    for step in range(sweep_steps):
        sel = np.arange(len(S)) % sweep_steps == step
        if not sel.any():
            continue
        s = S[sel]
        half = 1.959964 * s.std(ddof=1) / math.sqrt(len(s)) if len(s) > 1 else float("nan")
        lines.append(f"{G[sel][0]:.6g}, {s.mean():.6g}, {half:.3g}, " + ", ".join(f"{x:.6g}" for x in np.quantile(s, qs)) +
                     f", {len(s)}")
    return "\n".join(lines)


def sweep_table(columns, sweep_steps, stat=0, scale=1.0, labels=None, first_replication=0, name="stat"):
    """One row per point of a parameter sweep run as ONE ensemble (Program.sweep / cxxdes_b200_program.sweep_steps:
    replication r ran point (first_replication + r) % sweep_steps): mean of f64[stat] * scale over the
    replications of the point and the half-width of its 95 % interval."""
    x = np.asarray(columns.f64[stat], dtype=np.float64) * scale
    point = (first_replication + np.arange(len(x))) % sweep_steps
    lines = [f"point, {name}, ci95, replications"]
    for step in range(sweep_steps):
        v = x[point == step]
        if not len(v):
            continue
        half = 1.959964 * v.std(ddof=1) / math.sqrt(len(v)) if len(v) > 1 else float("nan")
        lines.append(f"{labels[step] if labels is not None else step}, {v.mean():.6g}, {half:.3g}, {len(v)}")
    return "\n".join(lines)


def summary(acc, names_f64=("f64[0]", "f64[1]"), names_i64=("i64[0]", "i64[1]"), run_ms=None):
    """A JSON-ready dict of one ensemble run: counts, checksum of checksums, and per statistic the ensemble mean
    with its 95 % interval (from the reduced accumulators alone)."""
    n = acc.n_replications
    out = {"n_replications": int(n), "n_events": int(acc.n_events), "n_errors": int(acc.n_errors),
           "trace_hash_sum": f"{acc.trace_hash_sum:016x}", "mean_end_time": acc.end_time_sum / n if n else float("nan")}
    for k, name in enumerate(names_f64):
        mean, half = mean_and_ci(acc, k)
        out[name] = {"mean": mean, "ci95": half}
    for k, name in enumerate(names_i64):
        out[name] = {"sum": int(acc.i64_sum[k]), "mean": acc.i64_sum[k] / n if n else float("nan")}
    if run_ms is not None:
        out["run_ms"] = float(run_ms)
        out["events_per_s"] = acc.n_events / (run_ms * 1e-3) if run_ms > 0 else float("nan")
    return out


def write_json(path, obj):
    import json
    with open(path, "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)
        f.write("\n")


def write_columns_csv(path, columns, names_f64=("f64_0", "f64_1"), names_i64=("i64_0", "i64_1")):
    """Per-replication results as CSV (one row per replication): what a user of the examples would otherwise
    collect by running the simulation object once per row."""
    hdr = ["replication", "end_time", "n_events", "trace_hash", "status", *names_f64, *names_i64]
    with open(path, "w") as f:
        f.write(", ".join(hdr) + "\n")
        for r in range(len(columns.end_time)):
            f.write(", ".join([str(r), str(int(columns.end_time[r])), str(int(columns.n_events[r])),
                               f"{int(columns.trace_hash[r]):016x}", str(int(columns.status[r])),
                               *[repr(float(columns.f64[k][r])) for k in range(abi.N_F64)],
                               *[str(int(columns.i64[k][r])) for k in range(abi.N_I64)]]) + "\n")


def time_series(ens, piece, n_pieces, stat=1, count=0):
    """Ensemble time series from run_for in pieces: after every piece of `piece` ticks the
    on-device reduction gives the cumulative sums; their differences are the per-interval statistics.

    Returns a list of dicts, one per interval: t_end (ticks), events (all replications), d_stat (increase of
    f64[stat] summed over replications, e.g. total latency), d_count (increase of i64[count], e.g. packets
    consumed), mean (= d_stat / d_count: e.g. the mean latency of the packets finished in the interval)."""
    rows = []
    prev_ev, prev_stat, prev_count = 0, 0.0, 0
    for j in range(1, n_pieces + 1):
        ens.run_for(piece)
        acc, _ = ens.reduce()
        ev, st, ct = acc.n_events, acc.f64_sum[stat], acc.i64_sum[count]
        rows.append({"t_end": j * piece, "events": ev - prev_ev, "d_stat": st - prev_stat, "d_count": ct - prev_count,
                     "mean": (st - prev_stat) / (ct - prev_count) if ct != prev_count else float("nan")})
        prev_ev, prev_stat, prev_count = ev, st, ct
    return rows


def batch_means(rows, warmup=1, z=1.959964):
    """Batch-means estimate from time_series() rows: the first `warmup` intervals are dropped (initial
    transient), the remaining interval means are treated as one sample each.  Returns (mean, half-width)."""
    xs = np.array([r["mean"] for r in rows[warmup:] if r["d_count"]], dtype=np.float64)
    if len(xs) < 2:
        return (float(xs[0]) if len(xs) else float("nan")), float("nan")
    return float(xs.mean()), float(z * xs.std(ddof=1) / math.sqrt(len(xs)))
