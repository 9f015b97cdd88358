#!/bin/bash
# round 2, call J (8 GPUs): bench --gpus 8
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2j_bench8.json 2> gpurun_out/r2j_bench8.err
echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2j_bench8.json') if x.startswith('{')]
j=json.loads(l[-1])
print(json.dumps({k:j[k] for k in ('value','e2e','extra','repeats') if k in j}, indent=1))
PY
tail -5 gpurun_out/r2j_bench8.err
