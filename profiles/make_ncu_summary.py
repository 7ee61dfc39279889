"""Turns one `ncu --set full` capture of the raster kernel into profiles/r1_raster_ncu_summary.json.
Usage: python profiles/make_ncu_summary.py gpurun_out/raster_r1_final.ncu-rep [cameras]"""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
cameras = int(sys.argv[2]) if len(sys.argv) > 2 else 1024 * 64
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def to_bytes(k):
    u, v = d[k]
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def to_ms(k):
    u, v = d[k]
    return float(v.replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]


traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
summary = {
    "source": "ncu --set full --clock-control none -k regex:raster_kernel -s 3 -c 1 python profiles/time_raster.py "
              f"({cameras} cameras of 64x64), one launch",
    "report": rep + " (scratch, not committed)",
    "kernel": d["Kernel Name"][1],
    "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
    "traffic_bytes_per_launch": traffic, "algorithmic_bytes_per_launch": cameras * 12 * 64 * 64,
    "duration_ms_under_ncu": to_ms("gpu__time_duration.sum"),
    "warp_instructions_per_camera": float(d["smsp__inst_executed.sum"][1].replace(",", "")) / cameras,
    "metrics": {k: {"unit": d[k][0], "value": d[k][1]} for k in keys if k in d},
}
json.dump(summary, open("profiles/r1_raster_ncu_summary.json", "w"), indent=1)
print(json.dumps({k: summary[k] for k in ("traffic_bytes_per_launch", "algorithmic_bytes_per_launch", "duration_ms_under_ncu",
                                          "warp_instructions_per_camera")}))
for k in keys[11:21]:
    if k in d:
        print(k, d[k])
