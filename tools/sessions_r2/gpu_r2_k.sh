#!/bin/bash
# round 2, call K: templated split + compat threshold: suite, all-shape timings, bench with natural leg
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -5 gpurun_out/r2k_pytest.log
NF=15 timeout 600 python tools/quick_bench.py C1 C2 C3 C4 C5 REFDEFAULT fast > gpurun_out/r2k_quick.log 2>&1
NF=15 SCREEN=0 timeout 600 python tools/quick_bench.py C1 C2 C3 C4 C5 REFDEFAULT fast >> gpurun_out/r2k_quick.log 2>&1
cat gpurun_out/r2k_quick.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
python - <<'PY'
import json
j=json.loads([x for x in open('gpurun_out/r2k_bench.json') if x.startswith('{')][-1])
print(j['value'], j['e2e']['value'], j['roofline']['kernel_ms_per_launch'], j['roofline']['hw_frac'], j['roofline']['frac'])
print(json.dumps(j['extra']['natural'], indent=1))
PY
