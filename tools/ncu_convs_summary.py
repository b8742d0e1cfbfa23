"""JSON summary (one record per launch) of an `ncu --set full` report of the conv kernels:
python tools/ncu_convs_summary.py gpurun_out/prof_convs_r1.ncu-rep > profiles/r1_ncu_full_convs_nusc18.json
bench.py reads `traffic` (dram read + write per launch) of the dominant conv from that file."""
import csv
import json
import re
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(h)}


def val(r, key, scale=1.0):
    if key not in col:
        return None
    try:
        v = float(r[col[key]].replace(",", ""))
    except ValueError:
        return None
    u = units[col[key]]
    if u in ("Mbyte", "MB"):
        v *= 1.0
    elif u in ("Kbyte", "KB"):
        v /= 1e3
    elif u in ("byte", "B"):
        v /= 1e6
    elif u in ("Gbyte", "GB"):
        v *= 1e3
    elif u == "ns":
        v /= 1e3
    elif u == "ms":
        v *= 1e3
    return v * scale


recs = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = re.sub(r"^void\s+", "", name).replace("<unnamed>::", "")
    short = re.sub(r"\(.*", "", short).replace("(int)", "").replace("(bool)", "")
    recs.append({
        "kernel": short,
        "us": val(r, "gpu__time_duration.sum"),
        "grid": re.sub(r"[(),]", " ", r[col["Grid Size"]]).split()[0],
        "dram_read_MB": val(r, "dram__bytes_read.sum"),
        "dram_write_MB": val(r, "dram__bytes_write.sum"),
        "tensor_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col
        else val(r, "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active"),
        "lts_pct": val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "issue_pct": val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active")
        if "sm__inst_issued.avg.pct_of_peak_sustained_active" in col
        else val(r, "smsp__issue_active.avg.pct"),
    })
json.dump(recs, sys.stdout, indent=0)
