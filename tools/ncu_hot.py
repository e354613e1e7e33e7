"""Top stalled SASS instructions of a kernel from an .ncu-rep (source page).  usage: ncu_hot.py rep [topN]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
# first kernel only
rows = list(csv.reader(out))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    data.append(dict(zip(hdr, r)))
tot = sum(int(d["# Samples"] or 0) for d in data)
print("instructions:", len(data), "samples:", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {c: sum(int(d[c] or 0) for d in data) for c in stall_cols}
print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for i, d in sorted(enumerate(data), key=lambda t: -int(t[1]["# Samples"] or 0))[:top]:
    s = int(d["# Samples"] or 0)
    main = sorted(((int(d[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {s:6d} {100.0 * s / max(tot, 1):5.1f}%  {d['Source'][:90]:90s} {main}")
