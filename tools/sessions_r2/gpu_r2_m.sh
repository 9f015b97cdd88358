#!/bin/bash
# round 2, call M (4 GPUs): bench --gpus 4 (both arms, as the driver's scaling run launches them)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 4 --steps 5 --warmup 2 > gpurun_out/r2m_ref4.json 2> gpurun_out/r2m_ref4.err
echo "ref rc=$?"; tail -c 400 gpurun_out/r2m_ref4.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2m_bench4.json 2> gpurun_out/r2m_bench4.err
echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2m_bench4.json') if x.startswith('{')]
j=json.loads(l[-1])
print(json.dumps({k:j[k] for k in ('value','e2e','extra','repeats') if k in j}, indent=1))
PY
tail -3 gpurun_out/r2m_bench4.err
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
