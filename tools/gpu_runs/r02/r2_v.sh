#!/bin/bash
# round 2, call V (gpurun --gpus 2): the driver's N>1 launch at HEAD — weak scaling line on 2 GPUs, both arms
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
echo "=== weak scaling line on $N GPUs"
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-library-baseline > gpurun_out/bench_weak_n$N.json 2> gpurun_out/bench_weak_n$N.err; cut -c1-700 gpurun_out/bench_weak_n$N.json; tail -2 gpurun_out/bench_weak_n$N.err
echo "=== reference arm under torchrun"
timeout 600 $TR bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/ref_n$N.json 2> gpurun_out/ref_n$N.err; cut -c1-300 gpurun_out/ref_n$N.json; tail -2 gpurun_out/ref_n$N.err
