#!/usr/bin/env python
"""Condense an `ncu --set full` report into the few numbers DESIGN.md / bench.py cite (kept under profiles/).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_v5_cfg3_summary.json [frames_in_capture]
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__instruction_throughput.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        k = {"kernel": d.get("Kernel Name")}
        for key in KEYS:
            if key in d:
                k[key] = {"value": d[key], "unit": u[key]}
        stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(d[h])
                  for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")}
        k["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        if frames:
            rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            if rd and wr:
                total = float(rd) * scale[u["dram__bytes_read.sum"]] + float(wr) * scale[u["dram__bytes_write.sum"]]
                k["frames_in_capture"] = frames
                k["dram_bytes_per_frame"] = total / frames
        res.append(k)
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res[0], indent=1)[:1500])


if __name__ == "__main__":
    main()
