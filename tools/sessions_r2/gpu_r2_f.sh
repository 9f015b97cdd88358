#!/bin/bash
# round 2, call F: new screen kernel: parity + timing
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_hardening_gpu.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -8 gpurun_out/r2f_pytest.log
NF=15 timeout 300 python tools/quick_bench.py C3 C4 C5 REFDEFAULT fast > gpurun_out/r2f_quick.log 2>&1
cat gpurun_out/r2f_quick.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<'PY'
import json
j=json.loads([x for x in open('gpurun_out/r2f_bench.json') if x.startswith('{')][-1])
print(j['value'], j['e2e']['value'], j['roofline']['kernel_ms_per_launch'], j['roofline']['hw_frac'], j['roofline']['frac'])
PY
