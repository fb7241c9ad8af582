"""The BASELINE.json configs as bench.py runs them: model parameters, clocks, sizes, how they shard, and the
roofline constants of their first-pass kernels.  Shared by bench.py, tools/gen_bench_gates.py and tools/bench_configs.py
(declarations only)."""
from libcxxdes_b200 import abi

SEED = 0x5EED
S_MS = ((1, abi.SECONDS), (1, abi.MILLISECONDS))
S_S = ((1, abi.SECONDS), (1, abi.SECONDS))


class Workload:
    def __init__(self, key, model, params, clock, replications, default_scaling, metric, describe, i_alg, i_alg_note,
                 gate_window):
        self.key = key
        self.model = model
        self.params = params
        self.clock = clock
        self.replications = replications          # weak: per GPU; strong: of the whole ensemble
        self.default_scaling = default_scaling
        self.metric = metric
        self.describe = describe
        self.i_alg = i_alg                        # algorithmic thread-instructions per event
        self.i_alg_note = i_alg_note
        self.gate_window = gate_window            # replications per committed parity-gate window


METRIC_MM1 = "simulated events/sec, 1M-replication M/M/1 ensemble at 1/2/4/8 B200"


def workloads(packets=10000, lam=5.0):
    return {
        "mm1": Workload(
            "mm1", abi.MODEL_MM1, abi.MM1Params(lam, 10.0, packets), S_MS, 1 << 20, "weak", METRIC_MM1,
            lambda r: (f"mm1 producer/consumer ensemble: {r} replications x {packets} packets, "
                       f"lambda={lam} mu=10.0, unit 1 s, precision 1 ms (BASELINE configs[1])"),
            82.0, "408 thread-instructions per packet pair / 5.0 events at rho = 0.5 (DESIGN.md 5.1)", 1024),
        "aloha": Workload(
            "aloha", abi.MODEL_ALOHA, abi.AlohaParams(64, 50, 0.1, 3.0, 1.0, 1000.0, 0), S_MS, 1 << 18, "weak",
            "simulated events/sec, ALOHA 64 stations, 256k replications sweeping offered load",
            lambda r: (f"examples/aloha.cpp: 64 stations, {r} replications, offered load G = 0.1 .. 3.0 in 50 steps "
                       "(replication id mod 50), G * 1000 packets per station (BASELINE configs[2])"),
            133.0, "sift-down 51 + sift-up 15 + digest 8 + station masks 15 + clock/check 4 + a variate (79) for every "
                   "second event 40 (DESIGN.md 5.3)", 256),
        "mmc": Workload(
            "mmc", abi.MODEL_MMC, abi.MMCParams(9.0, 1.0, 0.5, 8, 0, 1000), S_MS, 1 << 20, "weak",
            "simulated events/sec, M/M/c balking c=8, 1M replications x 1000 customers",
            lambda r: (f"M/M/c with any_of(acquire, timeout) balking (examples/resource.cpp pattern): c=8, lambda=9, mu=1, "
                       f"patience 0.5 s, 1000 customers, {r} replications (BASELINE configs[3])"),
            119.0, "sift-down 46 + one push per event 23 + digest 8 + customer word 10 + event logic 12 + clock/check 4 + "
                   "two variates per ~10 events 16 (DESIGN.md 5.3)", 1024),
        "archsim": Workload(
            "archsim", abi.MODEL_ARCHSIM, abi.ArchSimParams(1, 10, 16, 32, 1, 0, 1024), S_S, 1 << 22, "strong",
            "simulated events/sec, basic_arch_sim memory hierarchy, 4M replications sharded across the GPUs",
            lambda r: (f"examples/basic_arch_sim.cpp: mem2{{1,16}} -> mem1{{10}}, sequential driver, 1024 loads of "
                       f"philox % 32, {r} replications (BASELINE configs[4])"),
            53.0, "4-entry heap pop 10 + push 8 + digest 8 + state machine 15 + one LRU scan and one address draw per "
                  "load (5-10 events) 12 (DESIGN.md 5.3)", 1024),
        "archsim4": Workload(
            "archsim4", abi.MODEL_ARCHSIM, abi.ArchSimParams(1, 10, 16, 32, 4, 0, 256), S_S, 1 << 22, "strong",
            "simulated events/sec, basic_arch_sim memory hierarchy with 4 requesters, 4M replications",
            lambda r: (f"examples/basic_arch_sim.cpp with 4 requesters x 256 loads contending for the mutexes, "
                       f"{r} replications (SURVEY 8d config 5 variant)"),
            None, None, 1024),
    }


def gate_firsts(w):
    """Global ids at which a rank's shard can start: weak (rank * R) and strong (R split over 1, 2, 4, 8) scaling."""
    firsts = {r * w.replications for r in range(8)}
    for n in (1, 2, 4, 8):
        firsts |= {r * (w.replications // n) for r in range(n)}
    return sorted(firsts)
