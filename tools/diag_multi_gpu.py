#!/usr/bin/env python3
"""Per-GPU diagnosis of the M/M/1 kernel on a multi-GPU node (one process per GPU under torchrun):
kernel time, SM count, and the SM clock / power / throttle reasons sampled through NVML every 5 ms
WHILE the kernel runs.  usage: torchrun --nproc-per-node N tools/diag_multi_gpu.py"""
import os
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import pynvml  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import libcxxdes_b200 as L  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(local)
reps = int(os.environ.get("DIAG_REPLICATIONS", str(1 << 20)))
ens = L.Ensemble(L.MM1Params(5.0, 10.0, 10000), reps, seed=0x5EED, first_replication=rank * reps, device=local)
ens.time_unit(1, L.SECONDS).time_precision(1, L.MILLISECONDS)
ens.run().sync()
samples = []
stop = False


def pump():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h), pynvml.nvmlDeviceGetTemperature(h, 0)))
        time.sleep(0.005)


if world > 1:
    dist.barrier()
t = threading.Thread(target=pump, daemon=True)
t.start()
ms = []
for _ in range(6):
    ens.run().sync()
    ms.append(ens.last_run_ms())
stop = True
t.join()
clk = np.array([s[0] for s in samples]); pw = np.array([s[1] for s in samples])
reasons = 0
for s in samples:
    reasons |= s[2]
# two library kernels for comparison: are the slow GPUs slow on other compute-bound work too?
def timed(fn, n=5):
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
m1 = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); m2 = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
x = torch.rand(1 << 24, device="cuda") + 1.0
t_mm = timed(lambda: torch.matmul(m1, m2))
t_lg = timed(lambda: torch.lgamma(torch.lgamma(x) + 2.0))
props = torch.cuda.get_device_properties(local)
line = (f"gpu {local}: kernel ms {np.round(ms, 2).tolist()}  SMs {props.multi_processor_count}  sm clock MHz min/med/max "
        f"{clk.min()}/{int(np.median(clk))}/{clk.max()}  power W med/max {np.median(pw):.0f}/{pw.max():.0f}  temp max "
        f"{max(s[3] for s in samples)} C  reasons 0x{reasons:x}  samples {len(samples)}  | bf16 matmul 8192^3 {t_mm:.3f} ms "
        f"({2 * 8192 ** 3 / t_mm / 1e9:.0f} TF/s)  lgamma x2 on 16M floats {t_lg:.3f} ms")
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, line)
    if rank == 0:
        print("\n".join(out), flush=True)
    dist.destroy_process_group()
else:
    print(line, flush=True)
ens.close()
