#!/bin/bash
# round 2, call A: new live-reference tests + bench N=1
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_reference_live.py tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
