#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): BASELINE configs[3] (13B, global batch 32 over 8 GPUs, 256 new tokens), configs[4]
# (three-tower sweep on N GPUs), strong scaling of the headline job (global batch 8 split over N GPUs)
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$N" -ge 8 ]; then
echo "=== configs[3]: 13B, 4 sequences per GPU x $N GPUs, 256 new tokens"
timeout 600 $TR bench.py --gpus $N --preset 13b --batch 4 --new-tokens 256 --steps 2 --warmup 2 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_13b_n$N.json 2> gpurun_out/bench_13b_n$N.err; cut -c1-500 gpurun_out/bench_13b_n$N.json; tail -2 gpurun_out/bench_13b_n$N.err
fi
echo "=== configs[4]: tower sweep on $N GPUs"
timeout 600 $TR tools/bench_tower_sweep.py --batches 1,8,32,128,256 --reps 3 > gpurun_out/tower_sweep_n$N.jsonl 2> gpurun_out/tower_sweep_n$N.err; cut -c1-220 gpurun_out/tower_sweep_n$N.jsonl; tail -2 gpurun_out/tower_sweep_n$N.err
echo "=== strong scaling: global batch 8 over $N GPUs"
timeout 600 $TR bench.py --gpus $N --global-batch 8 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_strong_n$N.json 2> gpurun_out/bench_strong_n$N.err; cut -c1-400 gpurun_out/bench_strong_n$N.json; tail -2 gpurun_out/bench_strong_n$N.err
echo "=== done"
