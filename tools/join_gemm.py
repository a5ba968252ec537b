"""Joins an ncu launch list (one forward) with the recorded GEMM shapes, aggregates by shape."""
import collections, csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H = rows[hdr]; ki, vi = H.index("Kernel Name"), H.index("Metric Value")
times = [float(r[vi].replace(",", "")) / 1e3 for r in rows[hdr + 1:] if "k_gemm_tcgen05" in r[ki]]
shapes = json.load(open(sys.argv[2]))
assert len(times) == len(shapes), (len(times), len(shapes))
agg = collections.defaultdict(lambda: [0, 0.0])
for t, s in zip(times, shapes):
    agg[tuple(s)][0] += 1; agg[tuple(s)][1] += t
tot = sum(v[1] for v in agg.values())
print(f"GEMM total {tot:.1f} us over {len(times)} launches")
print(f"{'M':>6} {'N':>6} {'K':>6} {'bat':>4} taps flg {'n':>3} {'us':>9} {'us/call':>8} {'TFLOP/s':>8}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    M, N, K, b, taps, fl = k
    print(f"{M:6d} {N:6d} {K:6d} {b:4d} {taps:4d} {fl:3d} {v[0]:3d} {v[1]:9.1f} {v[1] / v[0]:8.1f} {2.0 * M * N * K * b * v[0] / v[1] / 1e6:8.1f}")
