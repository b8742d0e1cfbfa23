"""aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name"""
import collections, csv, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
recs = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for x in recs:
    name = re.sub(r'\(.*', '', x['Kernel Name'])[:64]
    v = float(x['Metric Value'].replace(',', ''))
    unit = x['Metric Unit']
    v = v / 1000 if unit == 'ns' else v * 1000 if unit == 'ms' else v
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"kernels {len(recs)}  total {tot:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v):9.1f} us {100*sum(v)/tot:5.1f}%  n={len(v):3d}  avg={sum(v)/len(v):8.1f}  max={max(v):8.1f}  {k}")
