"""Summarises an ncu gpu__time_duration launch list of tools/resnet_bench.py: the last training step, by kernel."""
import csv, collections, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/resnet_launches.csv'
rows=[r for r in csv.reader(open(path)) if len(r)>10 and r[0].isdigit()]
idx=[i for i,r in enumerate(rows) if 's2d_pack' in r[4]]
step=rows[idx[-1]:]
d=collections.defaultdict(list)
for r in step:
    name=re.sub(r'\(.*','',r[4])[:64]
    d[name].append(float(r[-1]))
tot=sum(sum(v) for v in d.values())
print("step launches", len(step), "total ms", round(tot/1e6,3))
for n,v in sorted(d.items(), key=lambda kv:-sum(kv[1]))[:20]: print(f"{sum(v)/1e6:8.3f} ms {100*sum(v)/tot:5.1f}% x{len(v):3d}  {n}")
if len(sys.argv) > 2:
    for r in step:
        if 'conv3d' in r[4] or 's2d_pack' in r[4]: print(f"{float(r[-1])/1e3:9.1f} us  {re.sub(r'\(.*','',r[4])[-36:]} grid {r[8]}")
