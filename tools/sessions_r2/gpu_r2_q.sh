#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -5 gpurun_out/r2q_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
python - <<'PY'
import json
j=json.loads([x for x in open('gpurun_out/r2q_bench.json') if x.startswith('{')][-1])
print(j['value'], j['e2e'], j['roofline']['kernel_ms_per_launch'], j['roofline']['hw_frac'], j['roofline']['frac'], j['impl_config'], j['gpu_launches'])
print(json.dumps(j['extra'], indent=0)[:1500])
PY
tail -3 gpurun_out/r2q_bench.err
