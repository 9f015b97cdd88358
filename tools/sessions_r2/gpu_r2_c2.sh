#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
NF=15 SCREEN=1 timeout 300 python tools/quick_bench.py REFDEFAULT fast > gpurun_out/r2c_quick.log 2>&1
NF=15 SCREEN=0 timeout 300 python tools/quick_bench.py REFDEFAULT fast >> gpurun_out/r2c_quick.log 2>&1
cat gpurun_out/r2c_quick.log
