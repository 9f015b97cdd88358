#!/bin/bash
# one 8-GPU session: multi-GPU tests, band benchmark (peer memory vs NCCL) at 4 and 8 ranks, frame-sharded bench at 8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -n 5
run_bands() { SD_BANDS_P2P=$2 SD_BANDS_VARIANT=$3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 tools/bench_bands.py 2>&1 | grep -E "^\{|Error|error" | tail -n 2; }
for N in 8 4; do
  run_bands $N 1 auto | tee -a gpurun_out/bands.jsonl
  run_bands $N 0 auto | tee -a gpurun_out/bands.jsonl
done
run_bands 8 1 fast | tee -a gpurun_out/bands.jsonl
python bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"; cat gpurun_out/bench_8gpu.json | cut -c1-900
python bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; cat gpurun_out/bench_2gpu.json | cut -c1-300
