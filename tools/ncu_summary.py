"""Condense an `ncu --page raw --csv` dump into the JSON summary committed under profiles/.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > gpurun_out/X_raw.csv
    python tools/ncu_summary.py gpurun_out/X_raw.csv profiles/X_summary.json

One object per profiled launch with the metrics DESIGN.md and bench.py's roofline.traffic quote.
"""
from __future__ import annotations

import csv
import json
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__cycles_active.avg",
    "sm__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second", "smsp__cycles_elapsed.avg.per_second",
]


def main(src: str, dst: str) -> None:
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = list(csv.reader(lines))
    header, units, body = rd[0], rd[1], rd[2:]
    col = {n: i for i, n in enumerate(header)}
    for r in body:
        if len(r) != len(header):
            continue
        o = {"Kernel Name": {"value": r[col["Kernel Name"]], "unit": ""},
             "Grid Size": {"value": r[col["Grid Size"]], "unit": ""},
             "Block Size": {"value": r[col["Block Size"]], "unit": ""}}
        for k in KEEP:
            if k in col:
                o[k] = {"value": r[col[k]], "unit": units[col[k]]}
        rows.append(o)
    with open(dst, "w") as f:
        json.dump(rows, f, indent=1)
    for o in rows:
        print(o["Kernel Name"]["value"][:60], o.get("gpu__time_duration.sum", {}).get("value"),
              o.get("dram__bytes_read.sum", {}).get("value"), o.get("dram__bytes_write.sum", {}).get("value"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
