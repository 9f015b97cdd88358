#!/bin/bash
cd "$GRAFT_REPO_ROOT/stereo_depth_b200/csrc"
for nb in 4 5 6; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall --fmad=false -DSD_SEC_MIN_BLOCKS=$nb -c secondary.cu -o build/secondary.o 2>/dev/null
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libstereo_b200.so build/*.o
  echo "SD_SEC_MIN_BLOCKS=$nb"
  (cd ../.. && NF=15 python tools/quick_bench.py C3 fast 2>&1 | tail -1)
done
