"""JSON summary (one record per launch) of an `ncu --set full` report of the memory-bound kernels:
    python tools/ncu_mem_summary.py gpurun_out/prof_mem.ncu-rep > profiles/r2_ncu_full_memory_kernels_nusc18.json
achieved DRAM GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum, against the measured
HBM copy bandwidth of MEASURED_PEAKS.json (6539.5 GB/s) and the nominal 8 TB/s (north_star)."""
import csv
import json
import os
import re
import subprocess
import sys

rep = sys.argv[1]
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(h)}


def val(r, key):
    if key not in col:
        return None
    try:
        v = float(r[col[key]].replace(",", ""))
    except ValueError:
        return None
    u = units[col[key]]
    scale = {"Kbyte": 1e3, "KB": 1e3, "Mbyte": 1e6, "MB": 1e6, "Gbyte": 1e9, "GB": 1e9, "byte": 1.0, "B": 1.0,
             "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    return v * scale


recs = []
for r in rows[2:]:
    name = re.sub(r"^void\s+", "", r[col["Kernel Name"]]).replace("<unnamed>::", "").replace("pn_detail::", "")
    name = re.sub(r"\(.*", "", name)
    us = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / us / 1e3 if us and rd is not None else None
    recs.append({
        "kernel": name, "us": us, "grid": re.sub(r"[(),]", " ", r[col["Grid Size"]]).split()[0],
        "dram_read_MB": rd / 1e6 if rd is not None else None, "dram_write_MB": wr / 1e6 if wr is not None else None,
        "dram_GBs": gbs, "pct_of_measured_hbm": 100 * gbs / hbm if gbs else None,
        "pct_of_nominal_8TBs": 100 * gbs / 8000 if gbs else None,
        "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "lts_pct": val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "regs": val(r, "launch__registers_per_thread"),
    })
json.dump(recs, sys.stdout, indent=0)
