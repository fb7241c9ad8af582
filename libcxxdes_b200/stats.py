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
