#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python tools/exp_host_chunks.py > gpurun_out/r2o_host_chunks.log 2>&1
cat gpurun_out/r2o_host_chunks.log
