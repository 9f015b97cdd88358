#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
SD_SCREEN_GROUPS=3 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest_g3.log 2>&1
echo "pytest (3 groups forced) rc=$?"; tail -4 gpurun_out/r2t_pytest_g3.log
for gr in 1 3; do
  echo "SD_SCREEN_GROUPS=$gr"
  SD_SCREEN_GROUPS=$gr NF=15 timeout 300 python tools/quick_bench.py C3 C5 C1 fast 2>&1 | grep -E "fast:" | sed -E 's/\(kernel-B.*per-frame/per-frame/'
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
python - <<'PY'
import json
j=json.loads([x for x in open('gpurun_out/r2t_bench.json') if x.startswith('{')][-1])
print(j['value'], j['e2e']['value'], j['roofline']['kernel_ms_per_launch'], j['roofline']['certified_screen']['evaluated_fraction'])
PY
