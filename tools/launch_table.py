"""Aggregates an ncu gpu__time_duration launch list by kernel for the LAST step of a run (from the last launch whose name matches
argv[2] on): python tools/launch_table.py launches.csv <first-kernel-substring> [--all]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:] if len(r) > vi and r[idi].isdigit()]
starts = [i for i, (k, v) in enumerate(data) if sys.argv[2] in k]
seg = data[starts[-1]:]
tot = sum(v for _, v in seg)
agg = collections.defaultdict(list)
for k, v in seg:
    name = re.sub(r"^void ", "", k)
    name = re.sub(r"\(.*", "", name).replace("mmad::", "")
    agg[name[:70]].append(v)
print(f"launches {len(seg)}  total {tot / 1e6:.3f} ms (ncu: cold-cache, serialised)")
for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:28]:
    print(f"{sum(v) / 1e6:8.3f} ms {100 * sum(v) / tot:5.1f}% x{len(v):3d}  {n}")
if "--all" in sys.argv:
    for k, v in seg:
        print(f"{v / 1e3:9.1f} us  {k[:100]}")
