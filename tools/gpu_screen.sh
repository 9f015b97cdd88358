#!/bin/bash
# screen bring-up: parity tests of the screened path, then quick timing with the screen on and off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "screen or full_size or fuzz or batches or adaptive" > gpurun_out/pytest_screen.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest_screen.log
SCREEN=1 NF=15 timeout 300 python tools/quick_bench.py C3 C5 C2 C4 C1 fast 2>&1 | tail -n 8
SCREEN=0 NF=15 timeout 300 python tools/quick_bench.py C3 fast 2>&1 | tail -n 3
