#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
j=json.loads([x for x in open('gpurun_out/bench.json') if x.startswith('{')][-1])
print(j['value'], j['e2e']['value'], j['e2e']['copy_ceiling_fps'], j['roofline']['kernel_ms_per_launch'], j['roofline']['hw_frac'], j['roofline']['frac'])
n=j['extra']['natural']; print({k:(v['fps'],v['ms_per_frame'],v.get('fps_screen_off'),v['evaluated_fraction']) for k,v in n.items() if isinstance(v,dict)})
print(j['extra']['c5'])
PY
