#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu shim"; timeout 600 python -m pytest tests/test_gpu_shim.py -m gpu -x -q 2>&1 | tail -5
for cfg in "ldg 0 0 1 0" "ldg 8 0 1 0" "ldg 0 0 1 1" "ldg 4 0 0 0" "tma 4 3 - -" "tma 4 2 - -" "tma 4 6 - -" "tma 2 3 - -" "tma 2 6 - -" "tma 8 3 - -"; do
set -- $cfg
if [ "$1" = "ldg" ]; then
  echo "=== skinny ldg warps=$2 mode=$4 db=$5"; VCB_SKINNY_WARPS=$2 VCB_SKINNY_MODE=$4 VCB_SKINNY_DB=$5 timeout 300 python tools/bench_skinny.py run 2>&1 | tail -1
else
  echo "=== skinny tma warps=$2 depth=$3"; VCB_SKINNY_IMPL=tma VCB_SKINNY_TMA_WARPS=$2 VCB_SKINNY_TMA_DEPTH=$3 timeout 300 python tools/bench_skinny.py run 2>&1 | tail -1
fi
done
echo "=== bench 2 GPUs (tma)"; VCB_SKINNY_IMPL=tma timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench6_2gpu.json 2> gpurun_out/bench6.err; tail -c 1500 gpurun_out/bench6_2gpu.json | cut -c1-600; tail -5 gpurun_out/bench6.err | cut -c1-300
