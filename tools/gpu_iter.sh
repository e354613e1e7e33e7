#!/usr/bin/env bash
# quick iteration on the Chamfer step: parity tests, per-kernel times (ncu launch list) and the bench line
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chamfer_gpu.py -x -q 2>&1 | grep -vE "^E   +\+" | tail -5
for v in ${VARIANTS:-0}; do
HP_RING_VARIANT=$v timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 60 --csv --log-file gpurun_out/l_$v.csv python tools/profile_chamfer.py 8 > /dev/null 2>&1
echo "variant $v (ncu, warm L2):"; python - "$v" <<'PY'
import csv, sys, collections
rows = list(csv.reader(open(f"gpurun_out/l_{sys.argv[1]}.csv")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) > mv: d[r[kn].split("(")[0][-40:]].append(float(r[mv].replace(",", "")))
for k, v in d.items():
    v = v[2:] if len(v) > 4 else v
    print(f"  {k:42s} n={len(v):3d} avg {sum(v)/len(v)/1e3:8.2f} us  min {min(v)/1e3:8.2f}")
PY
HP_RING_VARIANT=$v timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-other-paths --no-metrics-eval 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('  bench: ms/step %.5f value %.4e frac fwd %.4f fwd+bwd %.4f kernel_ms %.5f e2e %.4e clocks %s' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['fwd+bwd_frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['clocks']['sm_mhz']))"
done
