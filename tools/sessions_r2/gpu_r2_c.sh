#!/bin/bash
# round 2, call C: gather path tests + default-range timing
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_hardening_gpu.py tests/test_zz_reference_live.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -30 gpurun_out/r2c_pytest.log
NF=15 SCREEN=1 timeout 300 python tools/quick_bench.py REFDEFAULT C3 fast > gpurun_out/r2c_quick.log 2>&1
NF=15 SCREEN=0 timeout 300 python tools/quick_bench.py REFDEFAULT fast >> gpurun_out/r2c_quick.log 2>&1
cat gpurun_out/r2c_quick.log
