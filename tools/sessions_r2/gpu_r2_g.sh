#!/bin/bash
# round 2, call G: ncu of the screen kernel
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python tools/profile_run.py 3 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mbm_screen -s 1 -c 1 -o gpurun_out/r2_prof_screen -f python tools/profile_run.py 3 > gpurun_out/ncu_full_s.log 2>&1; echo "ncu full screen rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
