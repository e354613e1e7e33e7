#!/usr/bin/env bash
# per-kernel times (ncu launch list, warm L2) of the Chamfer step for several HP_RING_VARIANT values (timing experiments)
set -uo pipefail
mkdir -p gpurun_out
for v in ${VARIANTS:-0}; do
HP_RING_VARIANT=$v timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 60 --csv --log-file gpurun_out/l_$v.csv python tools/profile_chamfer.py 8 > /dev/null 2>&1
echo "variant $v (ncu, warm L2):"; python - "$v" <<'PY'
import csv, sys, collections
rows = list(csv.reader(open(f"gpurun_out/l_{sys.argv[1]}.csv")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) > mv and "hp::" in r[kn]: d[r[kn].split("(")[0][-40:]].append(float(r[mv].replace(",", "")))
for k, v in d.items():
    v = v[2:] if len(v) > 4 else v
    print(f"  {k:42s} n={len(v):3d} avg {sum(v)/len(v)/1e3:8.2f} us  min {min(v)/1e3:8.2f}")
PY
done
