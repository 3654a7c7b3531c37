#!/bin/bash
mkdir -p gpurun_out
echo "=== bench 2 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['config']['global_batch'], d['phases'])"; grep -v "^$" gpurun_out/bench_2gpu.err | tail -5 | cut -c1-200
echo "=== reference arm under torchrun (rank 0 only)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | cut -c1-300; echo "rc=$?"
