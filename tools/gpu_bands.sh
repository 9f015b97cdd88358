#!/bin/bash
# multi-GPU: NCCL + peer-memory band tests, then the C4 band benchmark both ways
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -n 15
for P in 0 1; do
SD_BANDS_P2P=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/bench_bands.py 2>&1 | grep -E "^\{|Error|error" | tail -n 3
done
