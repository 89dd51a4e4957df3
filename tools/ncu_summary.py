#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of one kernel launch into the JSON bench.py reads for roofline.traffic / ncu_pipes.

    python tools/ncu_summary.py gpurun_out/di_cfg5.ncu-rep profiles/ncu_di_kernel_cfg5.json --label "v4.3, cfg5, one B200"

Runs `ncu -i REP --page raw --csv` (works in the GPU-less container), takes the first launch matching --kernel, and keeps the
counters the judge asks for: DRAM bytes per launch, duration, L1/shared-memory data pipe, FP64 pipe, issue slots, occupancy, stall
reasons, executed instructions.  A text dump of the same counters is written next to the JSON."""
import argparse
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "sm__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "local_load_requests", "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum",
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out_json")
    ap.add_argument("--kernel", default="di_items_smem_kernel")
    ap.add_argument("--label", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    name_col = header.index("Kernel Name")
    launch = next((r for r in rows[2:] if a.kernel in r[name_col]), None)
    if launch is None:
        sys.exit(f"no launch of {a.kernel} in {a.report}")
    val, unit = {}, {}
    for h, u, v in zip(header, units, launch):
        if h in KEEP:
            try:
                val[h] = float(v.replace(",", ""))
            except ValueError:
                continue
            unit[h] = u
    def scaled(name):
        return val[name] * UNIT_SCALE.get(unit.get(name, ""), 1.0) if name in val else None
    traffic = (scaled("dram__bytes_read.sum") or 0.0) + (scaled("dram__bytes_write.sum") or 0.0)
    out = {
        "source": f"{Path(a.out_json).with_suffix('.txt').as_posix()} (ncu --set full --clock-control none of {a.kernel}; {a.label})",
        "kernel": launch[name_col][:120], "label": a.label,
        "traffic_bytes_per_launch": traffic, "duration_ms": scaled("gpu__time_duration.sum"),
        "pipes": {
            "l1tex_data_pipe_pct_of_peak": val.get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            "shared_memory_pct_of_peak": val.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
            "fp64_pipe_pct_of_peak": val.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": val.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": val.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": val.get("launch__registers_per_thread"),
            "lanes_per_instruction": val.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "l2_hit_pct": val.get("lts__t_sector_hit_rate.pct"),
        },
        "counters": {k: {"value": val[k], "unit": unit[k]} for k in KEEP if k in val},
    }
    Path(a.out_json).write_text(json.dumps(out, indent=1) + "\n")
    with open(Path(a.out_json).with_suffix(".txt"), "w") as f:
        f.write(f"# {a.label}\n# kernel: {launch[name_col]}\n")
        for k in KEEP:
            if k in val:
                f.write(f"{k} {val[k]} {unit[k]}\n")
    print(f"wrote {a.out_json}: traffic {traffic / 1e6:.2f} MB per launch, {out['duration_ms']:.2f} ms")


if __name__ == "__main__":
    main()
