#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -n 4
SCREEN=1 NF=15 timeout 300 python tools/quick_bench.py C1 C2 C3 C4 C5 fast 2>&1 | tail -n 5 | cut -c1-260
