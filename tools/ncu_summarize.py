"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (+ grid)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H = rows[hdr]
ki, vi, gi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size")
by = collections.defaultdict(lambda: [0, 0.0])
byg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    name = r[ki].split("(")[0]
    t = float(r[vi].replace(",", "")) / 1e3
    by[name][0] += 1; by[name][1] += t
    byg[(name, r[gi])][0] += 1; byg[(name, r[gi])][1] += t
tot = sum(v[1] for v in by.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in by.values())} launches")
for k, v in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  {k}")
if len(sys.argv) > 2:
    print("--- by grid ---")
    for k, v in sorted(byg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2])]:
        print(f"{v[1]:10.1f} us  n={v[0]:4d}  avg {v[1] / v[0]:8.1f} us  {k[0]} grid {k[1]}")
