#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2v_pytest.log
python tools/exp_natural_profile.py 2>&1 | tail -6
NF=15 timeout 300 python tools/quick_bench.py C3 C4 fast 2>&1 | grep "fast:" | sed -E 's/\(kernel-B.*per-frame/per-frame/'
NF=15 SCREEN=0 timeout 300 python tools/quick_bench.py C3 fast ws 2>&1 | grep -E "fast:|ws:" | sed -E 's/\(kernel-B.*per-frame/per-frame/'
