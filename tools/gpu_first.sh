#!/bin/bash
# first GPU bring-up: microbenchmark, golden fixtures from the reference, stage parity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
./tools/microbench/fadd_bench > gpurun_out/fadd_bench.txt 2>&1; echo "fadd_bench rc=$?"
tail -n 40 gpurun_out/fadd_bench.txt
python tests/golden/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?"; tail -n 12 gpurun_out/golden.log
python tools/first_parity.py generic > gpurun_out/first_parity.log 2>&1; echo "parity rc=$?"; tail -n 30 gpurun_out/first_parity.log
