#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
echo "== pytest gpu all"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== bench"; timeout 900 python bench.py --no-metrics-eval > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
print('value %.3e ms/step %.4f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))
for k,v in d['other_paths'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!='what'})
PY
tail -3 gpurun_out/bench.err
