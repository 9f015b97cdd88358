#!/bin/bash
# tests + smoke + bench + ncu launch list + one full ncu capture of the fused kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
python bench.py --impl reference --steps 3 --warmup 1 --frames 4 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?"; cat gpurun_out/bench_reference.json; tail -n 3 gpurun_out/bench_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
SCREEN=1 NF=30 timeout 300 python tools/quick_bench.py C3 fast 2>&1 | tail -n 1
SCREEN=1 NF=15 timeout 300 python tools/quick_bench.py C3 fast 2>&1 | tail -n 1
if [ "$1" == "ncu" ]; then
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
python tools/profile_run.py 3 > gpurun_out/plain2.log 2>&1 &&
SD_SCREEN=0 ncu --set full --clock-control none --import-source on -k regex:mbm_wta_fast -s 1 -c 1 -o gpurun_out/prof_kernelB python tools/profile_run.py 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full B (all levels) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mbm_screen -s 1 -c 1 -o gpurun_out/prof_screen python tools/profile_run.py 3 > gpurun_out/ncu_full_s.log 2>&1; echo "ncu full screen rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mbm_wta_fast -s 1 -c 1 -o gpurun_out/prof_kernelB_screened python tools/profile_run.py 3 > gpurun_out/ncu_full_f.log 2>&1; echo "ncu full B (screened) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"gray_pool|secondary|fill_kernel|pad_pooled" -s 5 -c 5 -o gpurun_out/prof_others python tools/profile_run.py 3 > gpurun_out/ncu_full2.log 2>&1; echo "ncu full others rc=$?"
fi
