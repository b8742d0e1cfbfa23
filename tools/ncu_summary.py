"""per-launch key counters from an ncu --set full report (raw page)"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
keep = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
keep = [k for k in keep if k in h]
idx = [h.index(k) for k in keep]
print(",".join(keep))
print(",".join(u[i] for i in idx))
for r in rows[2:]:
    print(",".join(r[i][:44].replace(",", ";") for i in idx))
