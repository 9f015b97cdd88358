#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
tail -12 gpurun_out/r2s_pytest.log
NF=15 timeout 300 python tools/quick_bench.py C3 C5 C4 fast 2>&1 | tail -3
SD_SEC_GROUP=0 NF=15 timeout 300 python tools/quick_bench.py C3 fast 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python - <<'PY'
import json
j=json.loads([x for x in open('gpurun_out/r2s_bench.json') if x.startswith('{')][-1])
print(j['value'], j['e2e']['value'], j['roofline']['kernel_ms_per_launch'], j['gpu_launches'])
n=j['extra']['natural']; print({k:(v['fps'],v.get('fps_screen_off')) for k,v in n.items() if isinstance(v,dict)})
PY
