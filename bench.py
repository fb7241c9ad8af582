#!/usr/bin/env python3
"""bench.py -- simulated events/s of the 1M-replication M/M/1 ensemble (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]              this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]        the reference CPU engine
  python bench.py --config {mm1,aloha,mmc,archsim,archsim4} ...    the other BASELINE configs, same one-line schema
  python bench.py --scaling {weak,strong} ...                      weak: --replications per GPU (global ids
                                                                   rank*R ... rank*R+R-1); strong: ONE ensemble of
                                                                   --replications split over the GPUs

One "step" = one pass of the hot path over one batch: every replication of the shard is run from its first event to
its last (environment::run), the per-replication results are reduced on the GPU and, at N > 1, the 88-byte accumulator
struct is all-reduced by the library's own collective (cxxdes_b200_reduce_allgpus: ncclAllReduce on a communicator
this script creates with ncclCommInitRank).  No data-path collective: replications share nothing.

Before anything is timed the results of one run are checked against checksums the CPU oracle produced offline
(tests/golden/mm1_config2_shards.json: every replication of every 1 Mi M/M/1 shard; tests/golden/bench_gates.json: the
first replications of every shard of the other configs): "parity_gate".  A mismatch aborts the bench.

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
sys.path.insert(0, str(ROOT / "tools"))

LAMBDA, MU = 5.0, 10.0
SEED = 0x5EED
I_ALG_SURVEY = 256.0     # SURVEY 8(d)'s estimate for M/M/1, made for draw-on-demand dispatch (one Philox block per draw)
BYTES_PER_REP = 60.0     # algorithmic HBM bytes per replication: the result row (7 x 8 B + 4 B)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="mm1", choices=["mm1", "aloha", "mmc", "archsim", "archsim4"],
                    help="which BASELINE config (default: configs[1], the headline M/M/1 ensemble)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="weak: --replications per GPU; strong: --replications in total, split over the GPUs "
                         "(default: weak, except archsim whose config is one 4 Mi ensemble sharded over the GPUs)")
    ap.add_argument("--replications", type=int, default=None, help="per GPU (weak) / in total (strong); default: the config's")
    ap.add_argument("--packets", type=int, default=10000)
    ap.add_argument("--no-parity-gate", action="store_true", help="skip the checksum comparison before timing")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel", default="ahead", choices=["ahead", "lockstep"],
                    help="first-pass M/M/1 kernel: generate-ahead (default) or the older draw-on-demand one (A/B)")
    ap.add_argument("--lam", type=float, default=LAMBDA, help="arrival rate (default = the headline 5.0)")
    a = ap.parse_args()
    import bench_workloads
    a.workload = bench_workloads.workloads(a.packets, a.lam)[a.config]
    if a.scaling is None:
        a.scaling = a.workload.default_scaling
    if a.replications is None:
        a.replications = a.workload.replications
    return a


def workload_name(a):
    how = "per GPU" if a.scaling == "weak" else "in total, split over the GPUs"
    return a.workload.describe(f"{a.replications} ({how})")


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


def cpu_run(oracle, w, n_reps, threads, first_rep=0):
    from libcxxdes_b200 import abi
    t0 = time.perf_counter()
    cols = oracle.run(w.model, w.params, abi.Clock.of(*w.clock), SEED, first_rep, n_reps, threads=threads)
    dt = time.perf_counter() - t0
    return int(cols.n_events.sum()), dt


def cpu_baseline(w, target_seconds):
    oracle, kind = load_cpu_oracle()
    threads = max(1, os.cpu_count() or 1)
    probe_reps = max(threads, 64)
    ev, dt = cpu_run(oracle, w, probe_reps, threads)
    rate = ev / dt
    reps = int(max(probe_reps, min(1 << 20, target_seconds * rate / max(1.0, ev / probe_reps))))
    ev, dt = cpu_run(oracle, w, reps, threads, first_rep=probe_reps)
    return {"value": ev / dt, "unit": "events/s", "cores": threads, "kind": kind,
            "sample": f"{reps} replications of the workload ({ev} events) in {dt:.2f} s, "
                      f"one replication per thread on {threads} host threads"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    oracle, kind = load_cpu_oracle()
    threads = max(1, os.cpu_count() or 1)
    w = a.workload
    # bounded sample per step: ~4 s of CPU work
    ev, dt = cpu_run(oracle, w, max(threads, 64), threads)
    per_rep = ev / max(threads, 64)
    reps = int(max(threads, min(a.replications, 4.0 * (ev / dt) / per_rep)))
    for _ in range(a.warmup):
        cpu_run(oracle, w, min(reps, 4 * threads), threads)
    total_ev, total_dt = 0, 0.0
    for s in range(a.steps):
        ev, dt = cpu_run(oracle, w, reps, threads, first_rep=s * reps)
        total_ev += ev
        total_dt += dt
    value = total_ev / total_dt
    sample = (f"{reps} replications of the workload per step on {threads} host threads "
              f"(one replication per thread, atomic work counter)")
    line = {
        "impl": "reference", "metric": w.metric, "value": value, "unit": "events/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total_dt / a.steps,
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "int64+f64",
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


def load_gates(a, first_rep, n_local, world, rank):
    """The committed checksums this rank's shard can be held against: (window, full) -- `window` = sums of its first K
    replications, `full` = sums over the whole shard (M/M/1 at the headline parameters only).  None = not available."""
    g = ROOT / "tests" / "golden"
    window = full = None
    w = a.workload
    standard = a.packets == 10000 and a.lam == LAMBDA
    try:
        gates = json.loads((g / "bench_gates.json").read_text())
        if standard and gates.get("seed") == SEED:
            window = gates["windows"].get(a.config, {}).get(str(first_rep))
            if window is not None and window["n"] > n_local:
                window = None
    except (OSError, ValueError):
        pass
    if a.config == "mm1" and standard:
        try:   # a whole 1 Mi shard (weak scaling, or one GPU)
            for sh in json.loads((g / "mm1_config2_shards.json").read_text())["shards"]:
                if sh["first_replication"] == first_rep and sh["n_replications"] == n_local:
                    full = sh
        except (OSError, ValueError, KeyError):
            pass
        if full is None:
            try:   # a strong-scaling shard of the first 1 Mi replications: whole blocks of the per-block sums
                blocks = json.loads((g / "mm1_config2_full.json").read_text())["blocks"]
                mine = [b for b in blocks if first_rep <= b["first"] and b["first"] + b["n"] <= first_rep + n_local]
                if mine and sum(b["n"] for b in mine) == n_local:
                    full = {k: sum(b[k] for b in mine) for k in ("n_events_sum", "end_time_sum", "trace_hash_sum",
                                                                 "packets_sum", "max_queue_sum")}
                    full["trace_hash_sum"] &= (1 << 64) - 1
            except (OSError, ValueError, KeyError):
                pass
    return window, full


def parity_gate(a, ens, host, first_rep, n_local, world, rank, dist, torch):
    """One run, fetched, held against the committed oracle checksums.  Returns the gate's description for the JSON line;
    raises SystemExit on a mismatch (a wrong digest must not print a throughput)."""
    if a.no_parity_gate:
        return "skipped (--no-parity-gate)"
    window, full = load_gates(a, first_rep, n_local, world, rank)
    ens.run()
    ens.fetch(host)
    problems, checked = [], []
    if window is not None:
        k = window["n"]
        got = {"n_events_sum": int(host.n_events[:k].sum(dtype=np.uint64)), "end_time_sum": int(host.end_time[:k].sum()),
               "trace_hash_sum": int(host.trace_hash[:k].sum(dtype=np.uint64))}
        for key, v in got.items():
            if v != window[key]:
                problems.append(f"rank {rank}: {key} of replications [{first_rep}, {first_rep + k}) is {v}, oracle {window[key]}")
        checked.append(f"first {k} replications of every shard")
    if full is not None:
        got = {"n_events_sum": int(host.n_events.sum(dtype=np.uint64)), "end_time_sum": int(host.end_time.sum()),
               "trace_hash_sum": int(host.trace_hash.sum(dtype=np.uint64)), "packets_sum": int(host.i64[0].sum()),
               "max_queue_sum": int(host.i64[1].sum())}
        for key, v in got.items():
            if v != full[key]:
                problems.append(f"rank {rank}: {key} over the shard at {first_rep} is {v}, oracle {full[key]}")
        checked.append(f"ALL {n_local} replications of every shard")
    if host.status.any():
        problems.append(f"rank {rank}: {int((host.status != 0).sum())} replications report a status")
    flags = torch.tensor([len(problems), 1 if window is not None else 0, 1 if full is not None else 0],
                         dtype=torch.int64, device="cuda")
    if world > 1:
        lo = flags.clone()
        dist.all_reduce(flags, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        have_window, have_full = int(lo[1]), int(lo[2])
    else:
        have_window, have_full = int(flags[1]), int(flags[2])
    if problems:
        print("PARITY GATE FAILED: " + "; ".join(problems), file=sys.stderr, flush=True)
    if int(flags[0]):
        raise SystemExit("bench.py: device results differ from the CPU oracle's checksums -- nothing was timed")
    parts = []
    if have_full:
        parts.append("every replication")
    if have_window:
        parts.append(f"first {a.workload.gate_window} replications of each shard")
    if not parts:
        return "no committed checksum for this shard layout (status words clean)"
    return ("ok: n_events / end_time / trace digest sums of " + " and of ".join(parts) +
            " equal the CPU oracle's (tests/golden; oracle/port, pinned to the reference engine)")


def run_b200(a):
    import torch
    import torch.distributed as dist

    import libcxxdes_b200 as L
    from libcxxdes_b200 import abi, nccl, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ensemble engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # the library's collective needs an ncclComm_t of ours: rank 0 draws the id, the process group (used for
        # nothing else but barriers and the bookkeeping of this script) hands the 128 bytes round
        uid = torch.zeros(nccl.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(nccl.get_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        comm = nccl.comm_init_rank(world, bytes(uid.cpu().numpy().tobytes()), rank)
        comm_ranks = nccl.comm_count(comm)
    else:
        comm_ranks = 1

    w = a.workload
    if a.scaling == "weak":
        first_rep, R = sharding.weak_shard_of(rank, a.replications)
        total_reps = a.replications * world
    else:
        first_rep, R = sharding.shard_of(rank, world, a.replications)
        total_reps = a.replications
    # a dedicated (non-default) torch stream: the library launches on it, so torch.cuda.Event sees the same stream
    # the kernels (and the library's NCCL call) run on
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    ens = L.Ensemble(w.params, R, seed=SEED, first_replication=first_rep, device=local_rank,
                     stream=stream.cuda_stream, flags=abi.FLAG_MM1_LOCKSTEP if a.kernel == "lockstep" else 0)
    ens.time_unit(*w.clock[0]).time_precision(*w.clock[1])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    host = abi.HostColumns(R, pinned=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        """run -> on-device reduce -> (N>1) the library's NCCL all-reduce of the accumulator struct."""
        flush.zero_()
        ens.run()
        if world > 1:
            ens.reduce_allgpus(comm, want_host=False)
        else:
            ens.reduce_async()

    gate = parity_gate(a, ens, host, first_rep, R, world, rank, dist, torch)
    kernel_name = ens.kernel_name()

    for _ in range(a.warmup):
        step()
    launches_before = ens.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()        # (before the barrier: forking nvidia-smi takes rank 0 ~15 ms, which the other ranks would
    barrier()                  #  otherwise spend waiting in the first all-reduce of their timed region)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        step()
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    last_kernel_ms = ens.last_run_ms()
    clocks = sampler.stop() if rank == 0 else None
    if clocks and world == 1:
        clocks.pop("per_gpu_sm_mhz", None)
    own_elapsed_ms = elapsed_ms
    launches = ens.launch_count() - launches_before
    passes = ens.pass_counts()

    # events processed in one step (every step simulates the same replications)
    acc, _ = ens.reduce()                      # this shard
    local_events = float(acc.n_events)
    if world > 1:
        totals = ens.reduce_allgpus(comm)      # the ensemble, through the library's collective
        events_per_step = float(totals.n_events)
        n_errors = int(totals.n_errors)
        assert totals.n_replications == total_reps, (totals.n_replications, total_reps)
        t = torch.tensor([elapsed_ms, float(launches), float(passes["second_pass"]), float(passes["helper"]),
                          float(passes["fallback"])], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        elapsed_ms = float(tmax[0])
        launches = int(t[1])
        passes = {"second_pass": int(t[2]), "helper": int(t[3]), "fallback": int(t[4])}
    else:
        events_per_step = local_events
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
    props = torch.cuda.get_device_properties(local_rank)
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    issue_peak = props.multi_processor_count * 4 * 32 * sm_max_mhz * 1e6
    # measured figures of the committed ncu capture: only if they were taken on THIS kernel instantiation
    traffic = measured_inst = None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        try:
            prof = json.loads(tp.read_text()).get("kernels", {}).get(kernel_name.split("+")[0])
            if prof:
                traffic = prof.get("dram_bytes_per_launch")
                measured_inst = prof.get("thread_inst_per_event")
        except (ValueError, OSError, AttributeError):
            pass
    i_alg = w.i_alg if a.kernel == "ahead" else None
    roofline = {
        "bound": "issue", "unit": "thread-inst/s", "peak": issue_peak, "traffic": traffic, "kernel": kernel_name,
        "kernel_ms": kernel_avg_ms,
        "measured_thread_inst_per_event": measured_inst,   # smsp__thread_inst_executed / events, committed ncu capture of this instantiation
        "peak_source": f"{props.multi_processor_count} SMs x 4 schedulers x 32 lanes x {sm_max_mhz:.0f} MHz, {peak_src}",
        "hbm": {"achieved": R * BYTES_PER_REP / (kernel_avg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": R * BYTES_PER_REP / (kernel_avg_ms * 1e-3) / 1e9 / hbm_peak,
                "alg_per_unit": f"{BYTES_PER_REP:.0f} B per replication (result row only; state stays on-chip)"},
    }
    if i_alg is not None:
        issue_achieved = local_events / (kernel_avg_ms * 1e-3) * i_alg
        roofline.update({"achieved": issue_achieved, "frac": issue_achieved / issue_peak,
                         "alg_per_unit": f"{i_alg:.0f} thread-instructions per simulated event: {w.i_alg_note}"})
        if a.config == "mm1":
            roofline["frac_with_survey_256"] = issue_achieved / i_alg * I_ALG_SURVEY / issue_peak
    else:
        roofline.update({"achieved": None, "frac": None, "alg_per_unit": None})
    if measured_inst:
        roofline["frac_executed"] = local_events / (kernel_avg_ms * 1e-3) * measured_inst / issue_peak

    # ---- end to end through the public API with HOST buffers ----
    params_bytes = C.sizeof(w.params)

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
           "d2h_bytes_per_step": host.nbytes() * world,
           "what": "steady-state run() + fetch() of every result column into pinned host memory per step; "
                   "create() (device allocation, parameter upload) is outside the timed region"}

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
        cpu = cpu_baseline(w, a.cpu_seconds)

    ens.close()
    if comm is not None:
        nccl.comm_destroy(comm)
    if rank == 0:
        how = (f"replications x{world} (weak: {a.replications} per GPU, global ids rank*R ...)" if a.scaling == "weak"
               else f"one ensemble of {a.replications} replications split over {world} GPUs (global ids [g*R/G, (g+1)*R/G))")
        line = {
            "metric": w.metric, "value": value, "unit": "events/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True,
            "scaling": a.scaling, "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "replications_total": total_reps,
                       "events_per_step": events_per_step, "sharding": how + ", no data-path collective; per step one "
                       "ncclAllReduce pair on the 88-byte accumulator struct, issued by cxxdes_b200_reduce_allgpus on a "
                       f"communicator of {comm_ranks} ranks" if world > 1 else how,
                       "l2": "flushed between steps (256 MiB memset inside the timed region)"},
            "parity_gate": gate,
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
            "n_replication_errors": n_errors, "last_kernel_ms": last_kernel_ms, "passes": passes,
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
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
