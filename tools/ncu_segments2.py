"""Attribute ncu warp-stall samples to the barrier-delimited segments of a kernel's SASS (first launch in the report).
usage: python tools/ncu_segments2.py report.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(h):
        continue
    if r[0] == "Address":
        break
    data.append(r)
ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
cols = {k: h.index(k) for k in ["stall_barrier", "stall_long_sb", "stall_math", "stall_no_inst", "stall_wait", "stall_short_sb",
                                "stall_not_selected", "stall_selected", "stall_mio", "stall_lg", "stall_dispatch", "stall_branch_resolving"]}


def new():
    return {"n": 0, "samp": 0, "ex": 0, "hmma": 0, "lds": 0, "sts": 0, "ldl": 0, "stl": 0, **{k: 0 for k in cols}}


seg, cur = [], new()
for r in data:
    s = r[ia]
    cur["n"] += 1
    cur["samp"] += int(r[isamp] or 0)
    cur["ex"] += int(r[iex] or 0)
    for k, i in cols.items():
        cur[k] += int(r[i] or 0)
    for key, pat in (("hmma", "HMMA"), ("lds", "LDS"), ("sts", "STS"), ("ldl", "LDL"), ("stl", "STL")):
        if pat in s:
            cur[key] += 1
    if "BAR.SYNC" in s:
        seg.append(cur)
        cur = new()
seg.append(cur)
tot = sum(x["samp"] for x in seg)
print("total samples", tot, "instructions", sum(x["n"] for x in seg))
for i, x in enumerate(seg):
    if x["samp"] > tot * 0.005:
        print(i, "n", x["n"], "hmma", x["hmma"], "lds", x["lds"], "sts", x["sts"], "ldl", x["ldl"], "stl", x["stl"], "samp%%%.1f" % (100 * x["samp"] / tot),
              " ".join("%s=%.1f" % (k[6:], 100 * x[k] / tot) for k in cols if x[k] > tot * 0.004))
