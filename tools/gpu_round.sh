#!/bin/bash
# The round's GPU check: tests + smoke + both bench arms; with "ncu": launch list + `--set full` captures of every kernel
# into gpurun_out/ (tools/make_profile_summaries.py turns them into the committed files under profiles/).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?"; tail -c 600 gpurun_out/bench_reference.json; tail -n 3 gpurun_out/bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
python bench.py --steps 2 --warmup 3 --no-extras --repeats 1 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras --repeats 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
python tools/profile_run.py 3 > gpurun_out/plain2.log 2>&1 &&
SD_SCREEN=0 ncu --set full --clock-control none --import-source on -k regex:mbm_wta_fast -s 1 -c 1 -f -o gpurun_out/prof_kernelB python tools/profile_run.py 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full B (all levels) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mbm_screen -s 1 -c 1 -f -o gpurun_out/prof_screen python tools/profile_run.py 3 > gpurun_out/ncu_full_s.log 2>&1; echo "ncu full screen rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mbm_wta_fast -s 1 -c 1 -f -o gpurun_out/prof_kernelB_screened python tools/profile_run.py 3 > gpurun_out/ncu_full_f.log 2>&1; echo "ncu full B (screened) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"gray_pool|secondary|fill_k|pad_pooled" -s 5 -c 5 -f -o gpurun_out/prof_others python tools/profile_run.py 3 > gpurun_out/ncu_full2.log 2>&1; echo "ncu full others rc=$?"
fi
