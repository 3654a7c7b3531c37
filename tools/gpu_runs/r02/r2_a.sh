#!/bin/bash
# round 2, call A: full GPU suite at HEAD (incl. the benched-shape tests), bench with library baseline, reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | sed -n 2p
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/pytest_r2a.log 2>&1; tail -25 gpurun_out/pytest_r2a.log
echo "=== bench"; timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; cut -c1-400 gpurun_out/bench_r2a.json; tail -3 gpurun_out/bench_r2a.err
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_r2a.json 2> gpurun_out/ref_r2a.err; cut -c1-1500 gpurun_out/ref_r2a.json; tail -2 gpurun_out/ref_r2a.err
echo "=== done"
