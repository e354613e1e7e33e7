#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chamfer_gpu.py -x -q -k "fused_step" 2>&1 | grep -vE "^E   +\+" | tail -15
timeout 300 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 80 --csv --log-file gpurun_out/l_step.csv python tools/profile_chamfer.py 8 step > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/l_step.csv")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) > mv and "hp::" in r[kn]: d[r[kn].split("(")[0][-40:]].append(float(r[mv].replace(",", "")))
for k, v in d.items():
    v = v[2:] if len(v) > 4 else v
    print(f"  {k:42s} n={len(v):3d} avg {sum(v)/len(v)/1e3:8.2f} us  min {min(v)/1e3:8.2f}")
PY
for pdl in 0 1; do
HP_NO_PDL=$pdl timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-other-paths --no-metrics-eval 2>gpurun_out/bench_iter.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('  HP_NO_PDL=$pdl bench: ms/step %.5f value %.4e frac fwd %.4f fwd+bwd %.4f kernel_ms %.5f e2e %.4e clocks %s' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['fwd+bwd_frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['clocks']['sm_mhz']))" || tail -5 gpurun_out/bench_iter.err
done
