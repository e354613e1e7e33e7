"""Aggregate stall samples by opcode for a SASS index range of a kernel in an .ncu-rep source page.
usage: ncu_loop.py rep start end"""
import csv, subprocess, sys, collections
rep, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi_]
data = []
for r in rows[hi_ + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    data.append(dict(zip(hdr, r)))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_all = sum(int(d["# Samples"] or 0) for d in data)
sel = data[lo:hi]
tot = sum(int(d["# Samples"] or 0) for d in sel)
print(f"range [{lo},{hi}): {tot} of {tot_all} samples ({100.0*tot/max(tot_all,1):.1f}%), instructions executed in range: {sum(int(d['Instructions Executed'] or 0) for d in sel)}")
byop = collections.defaultdict(lambda: collections.Counter())
cnt = collections.Counter()
for d in sel:
    op = d["Source"].split()[0]
    if op.startswith("@"):
        op = "@pred " + d["Source"].split()[1]
    op = op.split(".")[0]
    cnt[op] += 1
    for c in stall_cols:
        byop[op][c[6:]] += int(d[c] or 0)
for op, c in sorted(byop.items(), key=lambda kv: -sum(kv[1].values()))[:14]:
    s = sum(c.values())
    print(f"{op:14s} n={cnt[op]:4d} samples={s:5d} ({100.0*s/max(tot,1):4.1f}%)  " + ", ".join(f"{k}={v}" for k, v in c.most_common(5)))
agg = collections.Counter()
for c in byop.values():
    agg.update(c)
print("range stall totals:", ", ".join(f"{k}={v}" for k, v in agg.most_common(9)))
