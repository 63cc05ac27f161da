#!/usr/bin/env python3
"""Compact view of an ncu report: one line per kernel launch with the metrics the profiles/ summaries quote.

    ncu -i report.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_pick.py raw.csv [--json out.json]
"""
import csv
import json
import sys

PICK = [
    ("ms", "gpu__time_duration.sum"),
    ("dram_rd_GB", "dram__bytes_read.sum"),
    ("dram_wr_GB", "dram__bytes_write.sum"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("inst_M", "smsp__inst_executed.sum"),
    ("alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("regs", "launch__registers_per_thread"),
    ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
]
SCALE = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        rec = {"kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", "")}
        for name, metric in PICK:
            if metric not in col or r[col[metric]] == "":
                continue
            v = float(r[col[metric]].replace(",", ""))
            u = units[col[metric]]
            if name in ("ms", "dram_rd_GB", "dram_wr_GB") and u in SCALE:
                v *= SCALE[u]
            if name == "inst_M":
                v /= 1e6
            if name == "smem_dyn_KB" and u == "byte/block":
                v /= 1024
            rec[name] = round(v, 4)
        out.append(rec)
    for rec in out:
        print(json.dumps(rec))
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
