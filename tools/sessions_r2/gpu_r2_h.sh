#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python tools/exp_band_overhead.py > gpurun_out/r2h_band.log 2>&1
cat gpurun_out/r2h_band.log
