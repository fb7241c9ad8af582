#!/usr/bin/env python3
"""bench.py -- simulated events/s of the 1M-replication M/M/1 ensemble (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]              this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]        the reference CPU engine

One "step" = one pass of the hot path over one batch: every replication of the ensemble is run
from its first event to its last (environment::run), the per-replication results are reduced on
the GPU and, at N > 1, the accumulators are all-reduced over NCCL.  Weak scaling: every rank owns
`--replications` replications (global ids rank*R ... rank*R+R-1), no data-path collective.

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LAMBDA, MU = 5.0, 10.0
SEED = 0x5EED
I_ALG = 82.0             # algorithmic thread-instructions per event of the generate-ahead algorithm
I_ALG_SURVEY = 256.0     # SURVEY 8(d)'s estimate, made for draw-on-demand dispatch (one Philox block per draw)
BYTES_PER_REP = 60.0     # algorithmic HBM bytes per replication: the result row (7 x 8 B + 4 B)
CPU_LAMBDA = [LAMBDA]    # arrival rate of the CPU arm (follows --lam)
METRIC = "simulated events/sec, 1M-replication M/M/1 ensemble at 1/2/4/8 B200"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--replications", type=int, default=1 << 20, help="replications per GPU")
    ap.add_argument("--packets", type=int, default=10000)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel", default="ahead", choices=["ahead", "lockstep"],
                    help="first-pass M/M/1 kernel: generate-ahead (default) or the older draw-on-demand one (A/B)")
    ap.add_argument("--lam", type=float, default=LAMBDA, help="arrival rate (default = the headline 5.0)")
    return ap.parse_args()


def workload_name(a):
    return (f"mm1 producer/consumer ensemble: {a.replications} replications x {a.packets} packets per GPU, "
            f"lambda={a.lam} mu={MU}, unit 1 s, precision 1 ms (BASELINE configs[1])")


# ----------------------------------------------------------------------------------------------
# CPU side: the reference engine (oracle/_ref) when it was built, else the restated oracle.
# This is the one place bench.py executes anything under oracle/ (cpu_baseline / --impl reference).
# ----------------------------------------------------------------------------------------------
def load_cpu_oracle():
    sys.path.insert(0, str(ROOT / "tests"))
    import oracles
    ref = oracles.ref_oracle()
    if ref is not None:
        return ref, "reference"
    return oracles.port_oracle(), "port"


def cpu_run(oracle, n_reps, packets, threads, first_rep=0):
    import oracles
    from libcxxdes_b200 import abi
    p = abi.MM1Params(CPU_LAMBDA[0], MU, packets)
    t0 = time.perf_counter()
    cols = oracle.run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, SEED, first_rep, n_reps, threads=threads)
    dt = time.perf_counter() - t0
    return int(cols.n_events.sum()), dt


def cpu_baseline(packets, target_seconds):
    oracle, kind = load_cpu_oracle()
    threads = max(1, os.cpu_count() or 1)
    probe_reps = max(threads, 64)
    ev, dt = cpu_run(oracle, probe_reps, packets, threads)
    rate = ev / dt
    reps = int(max(probe_reps, min(1 << 20, target_seconds * rate / max(1.0, ev / probe_reps))))
    ev, dt = cpu_run(oracle, reps, packets, threads, first_rep=probe_reps)
    return {"value": ev / dt, "unit": "events/s", "cores": threads, "kind": kind,
            "sample": f"{reps} replications x {packets} packets ({ev} events) in {dt:.2f} s, "
                      f"one replication per thread on {threads} host threads"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    oracle, kind = load_cpu_oracle()
    threads = max(1, os.cpu_count() or 1)
    # bounded sample per step: ~4 s of CPU work
    ev, dt = cpu_run(oracle, max(threads, 64), a.packets, threads)
    per_rep = ev / max(threads, 64)
    reps = int(max(threads, min(a.replications, 4.0 * (ev / dt) / per_rep)))
    for _ in range(a.warmup):
        cpu_run(oracle, min(reps, 4 * threads), a.packets, threads)
    total_ev, total_dt = 0, 0.0
    for s in range(a.steps):
        ev, dt = cpu_run(oracle, reps, a.packets, threads, first_rep=s * reps)
        total_ev += ev
        total_dt += dt
    value = total_ev / total_dt
    sample = (f"{reps} replications x {a.packets} packets per step on {threads} host threads "
              f"(one replication per thread, atomic work counter)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total_dt / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64",
        "data": "synthetic", "config": {"workload": workload_name(a), "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU side
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            # no --id: one process watches every GPU of the node (per-GPU medians explain N > 1 runs)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        per_gpu = {}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                gpu = int(f[0])
                per_gpu.setdefault(gpu, []).append(float(f[1]))
                if gpu != self.index:
                    continue
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
               "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}
        if len(per_gpu) > 1:
            out["per_gpu_sm_mhz"] = [float(np.median(per_gpu[g])) for g in sorted(per_gpu)]
        return out


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def run_b200(a):
    import torch
    import torch.distributed as dist

    import libcxxdes_b200 as L
    from libcxxdes_b200 import abi, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ensemble engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    R = a.replications
    params = L.MM1Params(a.lam, MU, a.packets)
    # a dedicated (non-default) torch stream: the library launches on it, so torch.cuda.Event and
    # NCCL (torch.distributed) see the same stream the kernels run on
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    first_rep, _ = sharding.weak_shard_of(rank, R)
    ens = L.Ensemble(params, R, seed=SEED, first_replication=first_rep, device=local_rank,
                     stream=stream.cuda_stream, flags=abi.FLAG_MM1_LOCKSTEP if a.kernel == "lockstep" else 0)
    ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        """run -> on-device reduce -> (N>1) NCCL all-reduce of the accumulator struct."""
        flush.zero_()
        ens.run()
        dev_acc = ens.reduce_async()
        if world > 1:
            view = torch.as_tensor(sharding.DevicePointerView(dev_acc), device="cuda")
            sharding.all_reduce_words(view, dist)
        return dev_acc

    for _ in range(a.warmup):
        step()
    barrier()
    launches_before = ens.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record()
    for _ in range(a.steps):
        step()
        kernel_ms.append(None)
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    last_kernel_ms = ens.last_run_ms()
    clocks = sampler.stop() if rank == 0 else None
    if clocks and world == 1:
        clocks.pop("per_gpu_sm_mhz", None)
    own_elapsed_ms = elapsed_ms
    launches = ens.launch_count() - launches_before

    # events processed by this rank in one step (every step simulates the same replications)
    acc, _ = ens.reduce()
    if world > 1:
        # after the in-place all-reduce in step() the struct held global sums; reduce() recomputed local ones
        t = torch.tensor([elapsed_ms, float(acc.n_events), float(acc.n_errors), float(launches)],
                         dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        elapsed_ms = float(tmax[0])
        events_per_step = float(t[1])
        n_errors = int(t[2])
        launches = int(t[3])
    else:
        events_per_step = float(acc.n_events)
        n_errors = int(acc.n_errors)
    value = events_per_step * a.steps / (elapsed_ms * 1e-3)

    # ---- per-launch duration of the dominant kernel, CUDA events on the launching stream ----
    # (the library brackets its launches with its own events on the same stream)
    k_ms = []
    for _ in range(max(3, a.steps)):
        flush.zero_()
        ens.run()
        ens.sync()
        k_ms.append(ens.last_run_ms())
    kernel_avg_ms = float(np.mean(k_ms))
    local_events = float(acc.n_events)
    props = torch.cuda.get_device_properties(local_rank)
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    issue_peak = props.multi_processor_count * 4 * 32 * sm_max_mhz * 1e6
    issue_achieved = local_events / (kernel_avg_ms * 1e-3) * I_ALG
    traffic = None
    measured_inst = None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        try:
            prof = json.loads(tp.read_text())
            traffic = prof.get("mm1_ahead_dram_bytes_per_launch" if a.kernel == "ahead" else "mm1_fused_dram_bytes_per_launch")
            if a.kernel == "ahead":
                measured_inst = prof.get("mm1_ahead_thread_inst_per_event")
        except (ValueError, OSError):
            traffic = None
    roofline = {
        "bound": "issue", "achieved": issue_achieved, "peak": issue_peak, "unit": "thread-inst/s",
        "frac": issue_achieved / issue_peak, "traffic": traffic,
        "kernel": "mm1_ahead_kernel" if a.kernel == "ahead" else "mm1_fused_kernel<FAST>", "kernel_ms": kernel_avg_ms,
        "alg_per_unit": f"{I_ALG:.0f} thread-instructions per simulated event (DESIGN.md 5: 408 per packet pair / 5 events; "
                        f"SURVEY 8d estimated {I_ALG_SURVEY:.0f} for draw-on-demand dispatch, which this kernel undercuts)",
        "frac_with_survey_256": issue_achieved / I_ALG * I_ALG_SURVEY / issue_peak,
        "measured_thread_inst_per_event": measured_inst,   # smsp__thread_inst_executed / events, from the committed ncu capture
        "peak_source": f"{props.multi_processor_count} SMs x 4 schedulers x 32 lanes x {sm_max_mhz:.0f} MHz, {peak_src}",
        "hbm": {"achieved": R * BYTES_PER_REP / (kernel_avg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": R * BYTES_PER_REP / (kernel_avg_ms * 1e-3) / 1e9 / hbm_peak,
                "alg_per_unit": f"{BYTES_PER_REP:.0f} B per replication (result row only; state stays on-chip)"},
    }

    # ---- end to end through the public API with HOST buffers ----
    host = abi.HostColumns(R, pinned=True)
    params_bytes = C.sizeof(params)

    def e2e_step():
        # inputs: the model parameters travel host -> device with the launch; outputs: every
        # per-replication column device -> pinned host memory
        ens.run()
        ens.fetch(host)
        return int(host.n_events.sum())

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_events = 0
    for _ in range(a.steps):
        e2e_events += e2e_step()
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t[0])
        t = torch.tensor([float(e2e_events)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e2e_events = float(t[0])
    e2e = {"value": e2e_events / e2e_dt, "unit": "events/s", "h2d_bytes_per_step": params_bytes * world,
           "d2h_bytes_per_step": host.nbytes() * world}

    # per-rank view (which GPU, if any, holds the others back at N > 1)
    per_rank = None
    if world > 1:
        mine = torch.tensor([kernel_avg_ms, own_elapsed_ms / a.steps], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"kernel_ms": [round(float(t[0]), 3) for t in allr], "step_ms": [round(float(t[1]), 3) for t in allr]}
        if clocks and clocks.get("per_gpu_sm_mhz"):
            per_rank["sm_mhz"] = clocks.pop("per_gpu_sm_mhz")

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline(a.packets, a.cpu_seconds)

    ens.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "replications_total": R * world,
                       "events_per_step": events_per_step, "sharding": f"replications x{world}, no data-path collective; "
                       "NCCL all-reduce of the 88-byte accumulator struct per step",
                       "l2": "flushed between steps (256 MiB memset inside the timed region)"},
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
            "n_replication_errors": n_errors, "last_kernel_ms": last_kernel_ms,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if per_rank is not None:
            line["per_rank"] = per_rank
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    CPU_LAMBDA[0] = a.lam
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
