#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel launch) into a small text file for profiles/.

usage: tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.summary.txt [events_per_launch]
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "local_load", "smsp__inst_executed_op_local",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    events = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"== kernel: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for k in hdr:
            if any(k == w or k.startswith(w) for w in KEEP) or (k.startswith(STALL) and k.endswith("per_issue_active.ratio")):
                if d[k] not in ("", "0", "0.000000"):
                    lines.append(f"{k:95s} {d[k]:>22s} {u[k]}")
        if events:
            try:
                wi = float(d["smsp__inst_executed.sum"])
                if d.get("smsp__thread_inst_executed.sum", "") not in ("", "0"):
                    ti = float(d["smsp__thread_inst_executed.sum"])
                else:   # not collected by this set: warp instructions x average active lanes
                    ti = wi * float(d["smsp__thread_inst_executed_per_inst_executed.ratio"])
                ns = float(d["gpu__time_duration.sum"]) * {"ms": 1e6, "us": 1e3, "ns": 1.0, "s": 1e9}[u["gpu__time_duration.sum"]]
                lines.append(f"-- derived: {ti / events:.1f} thread-inst/event, {wi * 32 / events:.1f} issued lane-slots/event, "
                             f"{events / ns:.2f} Gevents/s under ncu (cold, serialised; do not quote as a bench number)")
            except (KeyError, ValueError):
                pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
