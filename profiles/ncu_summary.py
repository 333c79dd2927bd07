#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / bench.py cite.

    python profiles/ncu_summary.py gpurun_out/prof_r1b.ncu-rep > profiles/r01_step_stream_ncu.txt
"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "smsp__inst_executed_op_shared_ld.sum", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    print(f"# {rep}: {len(data)} launches of {data[0][name_col][:90]}")
    out = {}
    for m in WANT:
        if m in hdr:
            i = hdr.index(m)
            vals = [r[i] for r in data]
            print(f"{m:90s} {units[i]:12s} {'  '.join(vals)}")
            out[m] = vals
    rd = [float(v) for v in out["dram__bytes_read.sum"]]
    wr = [float(v) for v in out["dram__bytes_write.sum"]]
    unit = units[hdr.index("dram__bytes_read.sum")]
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[unit]
    print("# dram bytes per launch (read+write):", [(a + b) * scale for a, b in zip(rd, wr)])
    if len(sys.argv) > 2:
        json.dump({"dram_bytes_per_launch": sum((a + b) * scale for a, b in zip(rd, wr)) / len(rd),
                   "boards_per_launch": int(sys.argv[3]) if len(sys.argv) > 3 else 67108864,
                   "source": rep}, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
