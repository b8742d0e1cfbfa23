"""hottest SASS instructions (warp-stall samples) of one kernel in an ncu report:
python tools/ncu_hot.py <report.ncu-rep> <kernel name> [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
ci = {k: i for i, k in enumerate(h)}
i_s, i_src = ci["# Samples"], ci["Source"]
data = []
for n, r in enumerate(rows[hdr + 1:]):
    if len(r) <= max(i_s, i_src) or r[0] == "Address" or r[0] == "Kernel Name":
        break
    try:
        data.append((float(r[i_s] or 0), n, r[i_src]))
    except ValueError:
        break
tot = sum(x for x, _, _ in data) or 1
print(f"{kern}: {len(data)} instructions, {int(tot)} samples")
for x, n, src in sorted(data, reverse=True)[:top]:
    print("%5.1f%%  #%4d  %s" % (100 * x / tot, n, src.strip()[:140]))
