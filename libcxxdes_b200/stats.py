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
