#!/usr/bin/env python3
"""Reads an `ncu --set full` report of a bench.py run and records per-kernel DRAM bytes per launch in profiles/traffic.json.

    python tools/ncu_traffic.py <report.ncu-rep> <workload key, e.g. decode64k/log/65536/1073741824/L3>

bench.py's roofline.traffic reads that file (the number is per launch, like roofline.achieved).
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                          "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    name_i = hdr.index("Kernel Name")
    rd_i, wr_i, t_i = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    units = rows[1]

    def scale(u):
        return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1, "ms": 1,
                "nsecond": 1e-6, "second": 1e3}.get(u, 1)
    rec = {}
    for r in rows[2:]:
        if len(r) <= max(rd_i, wr_i):
            continue
        name = r[name_i].split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()
        rd = float(r[rd_i].replace(",", "")) * scale(units[rd_i])
        wr = float(r[wr_i].replace(",", "")) * scale(units[wr_i])
        ms = float(r[t_i].replace(",", "")) * scale(units[t_i])
        if rd != rd or wr != wr:      # ncu could not collect the counters for this launch
            continue
        rec[name] = {"dram_bytes": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr), "ncu_ms": round(ms, 4)}   # last launch wins
    path = os.path.join(ROOT, "profiles", "traffic.json")
    allrec = json.load(open(path)) if os.path.exists(path) else {}
    allrec[key] = rec
    allrec[key]["_source"] = os.path.basename(rep)
    json.dump(allrec, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
