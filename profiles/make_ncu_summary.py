"""Turns `ncu --set full` captures into the committed summaries under profiles/.

  python profiles/make_ncu_summary.py raster  gpurun_out/raster_X.ncu-rep  profiles/r2_raster_ncu_summary.json [cameras] [res]
  python profiles/make_ncu_summary.py kernels gpurun_out/Y.ncu-rep        profiles/r2_Y_ncu_summary.json

`raster`: one launch of raster_kernel (the dominant kernel): DRAM traffic against the algorithmic bytes, warp
instructions per camera, pipe / issue / stall figures and the executed-instruction footprint.
`raster2`: the two-pass 64x64 raster (draw pass raster_kernel<..., 1> + raster_finish_kernel, captured back to back):
the same figures per pass, and their sum against the algorithmic bytes.
`kernels`: every launch in the report, one row per kernel name (the first launch of each), same figures.
"""
import csv
import json
import subprocess
import sys

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [{h: (u, v) for h, u, v in zip(hdr, units, r)} for r in rows[2:]]


def num(d, k):
    u, v = d[k]
    return float(v.replace(",", "")) * SCALE.get(u, 1.0)


def footprint(rep):
    """SASS instructions of the kernel, those executed at all, and those executed at least 0.5 times per 1000 warps of work
    (the instruction-cache footprint of the steady state)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, ie = None, []
    for r in rows:
        if r and r[0] == "Address":
            if hdr is not None:
                break               # first kernel only
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            try:
                ie.append(float(r[hdr.index("Instructions Executed")].replace(",", "")))
            except ValueError:
                pass
    return ie


def summarize(d):
    s = {"kernel": d["Kernel Name"][1], "duration_ms_under_ncu": num(d, "gpu__time_duration.sum"),
         "dram_bytes_read": num(d, "dram__bytes_read.sum"), "dram_bytes_write": num(d, "dram__bytes_write.sum"),
         "warp_instructions": num(d, "smsp__inst_executed.sum"),
         "metrics": {k: {"unit": d[k][0], "value": d[k][1]} for k in KEYS if k in d}}
    s["traffic_bytes_per_launch"] = s["dram_bytes_read"] + s["dram_bytes_write"]
    return s


def main():
    mode, rep, out = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = rows_of(rep)
    if mode == "raster":
        cameras = int(sys.argv[4]) if len(sys.argv) > 4 else 1024 * 64
        res = int(sys.argv[5]) if len(sys.argv) > 5 else 64
        d = [r for r in rows if "raster_kernel" in r["Kernel Name"][1]][0]
        s = summarize(d)
        s["source"] = (f"ncu --set full --clock-control none --import-source on -k regex:raster_kernel -c 1 python profiles/time_raster.py "
                       f"({cameras} cameras of {res}x{res}), one launch")
        s["report"] = rep + " (scratch, not committed)"
        s["algorithmic_bytes_per_launch"] = cameras * 12 * res * res
        s["traffic_over_algorithmic"] = s["traffic_bytes_per_launch"] / s["algorithmic_bytes_per_launch"]
        s["warp_instructions_per_camera"] = s["warp_instructions"] / cameras
        ie = footprint(rep)
        if ie:
            s["sass_instructions"] = len(ie)
            s["sass_instructions_executed"] = sum(1 for x in ie if x > 0)
            s["sass_instructions_executed_at_least_half_per_camera"] = sum(1 for x in ie if x >= 0.5 * cameras)
        json.dump(s, open(out, "w"), indent=1)
        print(json.dumps({k: s[k] for k in ("kernel", "duration_ms_under_ncu", "traffic_over_algorithmic", "warp_instructions_per_camera")}))
    elif mode == "raster2":
        cameras = int(sys.argv[4]) if len(sys.argv) > 4 else 1024 * 64
        res = int(sys.argv[5]) if len(sys.argv) > 5 else 64
        draw = [r for r in rows if "raster_kernel" in r["Kernel Name"][1] and "1, 1>" in r["Kernel Name"][1]][0]
        fin = [r for r in rows if "raster_finish_kernel" in r["Kernel Name"][1]][0]
        passes = [summarize(draw), summarize(fin)]
        for q in passes:
            q["warp_instructions_per_camera"] = q["warp_instructions"] / cameras
        s = {"kernel": "raster, 64x64 two-pass form: " + " + ".join(q["kernel"].split("(")[0] for q in passes),
             "duration_ms_under_ncu": sum(q["duration_ms_under_ncu"] for q in passes),
             "dram_bytes_read": sum(q["dram_bytes_read"] for q in passes), "dram_bytes_write": sum(q["dram_bytes_write"] for q in passes),
             "warp_instructions": sum(q["warp_instructions"] for q in passes), "passes": passes}
        s["traffic_bytes_per_launch"] = s["dram_bytes_read"] + s["dram_bytes_write"]
        s["source"] = (f"ncu --set full --clock-control none -k regex:raster -c 2 python profiles/time_raster.py "
                       f"({cameras} cameras of {res}x{res}), one launch of each pass")
        s["report"] = rep + " (scratch, not committed)"
        s["algorithmic_bytes_per_launch"] = cameras * 12 * res * res
        s["traffic_over_algorithmic"] = s["traffic_bytes_per_launch"] / s["algorithmic_bytes_per_launch"]
        s["warp_instructions_per_camera"] = s["warp_instructions"] / cameras
        s["note"] = ("the traffic above the algorithmic bytes is the hand-over between the passes (bitplanes and lists of "
                     "border-crossing faces, written by the draw pass and read by the finish pass)")
        json.dump(s, open(out, "w"), indent=1)
        print(json.dumps({k: s[k] for k in ("kernel", "duration_ms_under_ncu", "traffic_over_algorithmic", "warp_instructions_per_camera")}))
    else:
        seen, res = set(), []
        for d in rows:
            name = d["Kernel Name"][1]
            if name in seen:
                continue
            seen.add(name)
            res.append(summarize(d))
        json.dump({"report": rep + " (scratch, not committed)", "kernels": res}, open(out, "w"), indent=1)
        for s in res:
            m = s["metrics"]
            print(s["kernel"][:60], f"{s['duration_ms_under_ncu']:.3f} ms", "issue", m["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"],
                  "lanes", m["smsp__thread_inst_executed_per_inst_executed.ratio"]["value"],
                  "fma", m["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]["value"])


if __name__ == "__main__":
    main()
