#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
F=8 timeout 300 python tools/exp_overlap_sc.py > gpurun_out/r2l_overlap.log 2>&1
F=4 timeout 300 python tools/exp_overlap_sc.py >> gpurun_out/r2l_overlap.log 2>&1
cat gpurun_out/r2l_overlap.log
NF=15 timeout 300 python tools/quick_bench.py C3 fast 2>&1 | tail -1
