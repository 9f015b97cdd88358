#!/bin/bash
# round 2, call B: whole GPU suite with durations
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -40 gpurun_out/r2b_pytest.log
