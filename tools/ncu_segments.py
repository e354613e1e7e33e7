"""Samples per barrier-delimited SASS segment of the first kernel in an .ncu-rep.  usage: ncu_segments.py rep"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi_]
data = []
for r in rows[hi_ + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    data.append(dict(zip(hdr, r)))
tot = sum(int(d["# Samples"] or 0) for d in data)
seg, segs = [], []
for i, d in enumerate(data):
    seg.append((i, d))
    if d["Source"].lstrip().startswith("BAR") or i == len(data) - 1:
        segs.append(seg)
        seg = []
print("total samples", tot, "segments", len(segs))
for sg in segs:
    s = sum(int(d["# Samples"] or 0) for _, d in sg)
    if s < 0.01 * tot:
        continue
    ops = collections.Counter(d["Source"].split()[0].split(".")[0] if not d["Source"].lstrip().startswith("@") else d["Source"].split()[1].split(".")[0] for _, d in sg)
    ex = sum(int(d["Instructions Executed"] or 0) for _, d in sg)
    stall = collections.Counter()
    for _, d in sg:
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                stall[h[6:]] += int(d[h] or 0)
    print(f"[{sg[0][0]:5d},{sg[-1][0]:5d}] {100.0*s/tot:5.1f}%  inst_exec={ex:9d}  top ops: {dict(ops.most_common(4))}  stalls: {dict(stall.most_common(4))}")
