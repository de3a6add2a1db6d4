#!/usr/bin/env python
"""Reads an .ncu-rep (one `ncu --set full` capture brought back from the GPU box) and writes the two summaries that
profiles/ keeps: <out>_key_metrics.json (DRAM bytes, duration, pipe utilisation, occupancy — `--page raw`) and
<out>_full_details.txt (`--page details`).  Usage: python scripts/ncu_extract.py gpurun_out/x.ncu-rep profiles/r2_x_ncu"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next((i for i, r in enumerate(rows) if r and r[0] == "ID"), None)
    res = {}
    if hdr is not None and len(rows) > hdr + 2:
        names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
        res["kernel"] = vals[names.index("Kernel Name")] if "Kernel Name" in names else ""
        for k in KEYS:
            if k in names:
                i = names.index(k)
                res[k] = [vals[i], units[i]]
    with open(out + "_key_metrics.json", "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    with open(out + "_full_details.txt", "w") as f:
        f.write(det)
    print(out, res.get("kernel", "?"), res.get("gpu__time_duration.sum"), res.get("dram__bytes_read.sum"))


if __name__ == "__main__":
    main()
