"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
agg = collections.defaultdict(lambda: [0, 0.0])
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 0
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    n += 1
    if n <= skip:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
    name = r[ki].split("(")[0].replace("void ", "").replace("diqt::", "")
    key = name[:48] + (" " + r[gi] if len(sys.argv) > 3 else "")
    agg[key][0] += 1
    agg[key][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{n - skip} launches, {tot:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:70s} {v[0]:5d} {v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  avg {v[1]/v[0]:7.2f}")
