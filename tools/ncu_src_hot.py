"""Condenses `ncu --page source --csv` (SASS view; second row is the header) into per-opcode totals
and a linear listing with instruction counts, so hot loops can be read off.
usage: zcat x.csv.gz | python tools/ncu_src_hot.py [min_inst_for_listing]"""
import csv, sys, collections
rows = list(csv.reader(sys.stdin))
H = rows[1]
col = {h: i for i, h in enumerate(H)}
i_src, i_inst, i_samp, i_thr = col["Source"], col["Instructions Executed"], col["# Samples"], col["Avg. Predicated-On Threads Executed"]
thresh = int(sys.argv[1]) if len(sys.argv) > 1 else 0
items = []
for r in rows[2:]:
    if len(r) <= i_inst or not r[i_inst].replace(",", "").isdigit():
        continue
    items.append((int(r[i_inst] or 0), int(r[i_samp] or 0), r[i_src].strip(), r[i_thr]))
tot_i, tot_s = sum(x[0] for x in items), sum(x[1] for x in items)
print(rows[0][1][:80])
print(f"total inst {tot_i} samples {tot_s} sass rows {len(items)}")
ops, ops_s = collections.Counter(), collections.Counter()
for ni, ns, src, _ in items:
    t = src.split()
    op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
    op = op.split(".")[0]
    ops[op] += ni; ops_s[op] += ns
print("--- by opcode ---")
for op, ni in ops.most_common(28):
    print(f"{ni:12d} {100*ni/max(1,tot_i):5.1f}%  samples {100*ops_s[op]/max(1,tot_s):5.1f}%  {op}")
print("--- listing (inst >= thresh) : idx inst samples avg_threads sass ---")
for k, (ni, ns, src, th) in enumerate(items):
    if ni >= thresh:
        print(f"{k:5d} {ni:11d} {ns:6d} {th:>5s}  {src[:100]}")
