#!/bin/bash
# Suggested FIRST gpurun call of round 2 (≈6 min): re-validate HEAD, then the two measurements round 1 left open —
# the HF-fp16 library baseline with token comparison, and a 2-GPU run of the bench (use `gpurun --gpus 2` for the last part).
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/pytest_r2.log 2>&1; tail -4 gpurun_out/pytest_r2.log
echo "=== bench"; timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; cut -c1-300 gpurun_out/bench_r2.json; tail -2 gpurun_out/bench_r2.err
echo "=== HF library baseline, tiny (fp16, sanity) then 7B with token comparison"
timeout 120 python tools/bench_hf_library.py --preset tiny --batch 2 --new-tokens 8 --reps 1 --compare-engine 2>&1 | tail -2
timeout 400 python tools/bench_hf_library.py --preset 7b --batch 8 --new-tokens 128 --reps 1 --compare-engine > gpurun_out/hf_library_r2.json 2> gpurun_out/hf_library_r2.err; cut -c1-600 gpurun_out/hf_library_r2.json; tail -3 gpurun_out/hf_library_r2.err
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "=== 2-GPU bench"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; cut -c1-300 gpurun_out/bench_r2_n2.json; tail -2 gpurun_out/bench_r2_n2.err
fi
echo "=== done"
