#!/bin/bash
mkdir -p gpurun_out
python tools/profile_run.py 3 > gpurun_out/plain2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/screen_launches.csv python tools/profile_run.py 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mbm_screen -s 1 -c 1 -o gpurun_out/prof_screen python tools/profile_run.py 3 > gpurun_out/ncu_full_s.log 2>&1; echo "ncu full screen rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mbm_wta_fast -s 1 -c 1 -o gpurun_out/prof_fast_screened python tools/profile_run.py 3 > gpurun_out/ncu_full_f.log 2>&1; echo "ncu full fast rc=$?"
