#!/bin/bash
# round 2, call E (2 GPUs): multi-GPU tests + bench --gpus 2 (bands_c4 / c5 legs)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_zz_reference_live.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -8 gpurun_out/r2e_pytest.log
timeout 900 python bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_bench2.json 2> gpurun_out/r2e_bench2.err
echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2e_bench2.json') if x.startswith('{')]
j=json.loads(l[-1])
print(json.dumps({k:j[k] for k in ('value','e2e','extra','repeats') if k in j}, indent=1))
PY
tail -5 gpurun_out/r2e_bench2.err
