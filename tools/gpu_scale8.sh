#!/bin/bash
# one 8-GPU session: multi-GPU tests, band benchmark (peer memory vs NCCL) at 8 ranks, frame-sharded bench at 8 and 4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -n 3
run_bands() { SD_BANDS_P2P=$2 SD_BANDS_VARIANT=$3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 tools/bench_bands.py 2>&1 | grep -E "^\{|Error|error" | tail -n 2; }
rm -f gpurun_out/bands.jsonl
run_bands 8 1 auto | tee -a gpurun_out/bands.jsonl
run_bands 8 0 auto | tee -a gpurun_out/bands.jsonl
python bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/bench_8gpu.err | grep "^{" > gpurun_out/bench_8gpu.json; echo "bench8 rc=$?"; cut -c1-200 gpurun_out/bench_8gpu.json
python bench.py --gpus 4 --steps 5 --warmup 3 2> gpurun_out/bench_4gpu.err | grep "^{" > gpurun_out/bench_4gpu.json; echo "bench4 rc=$?"; cut -c1-200 gpurun_out/bench_4gpu.json
