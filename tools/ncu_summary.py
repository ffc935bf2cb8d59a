#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) per kernel: duration, DRAM bytes, pipe/issue utilisation,
occupancy and the top warp-stall reasons.   python tools/ncu_summary.py file.ncu-rep [--json out.json]"""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__cycles_active.avg", "smsp__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        o = {"kernel": d["Kernel Name"][:80]}
        for k in KEYS:
            if k in d:
                o[k] = f"{d[k]} {u[k]}".strip()
        stalls = []
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[k]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        o["top_stalls_warps_per_issue"] = {n: round(v, 3) for v, n in stalls[:6]}
        res.append(o)
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(res, f, indent=1)
    for o in res:
        print(json.dumps(o, indent=1))


if __name__ == "__main__":
    main()
