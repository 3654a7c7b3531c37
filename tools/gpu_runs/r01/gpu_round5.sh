#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu (default)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== pytest -m gpu (tc5 attention + tma skinny)"; VCB_VIT_ATTN=tc5 VCB_SKINNY_IMPL=tma timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== vit tc5"; VCB_VIT_ATTN=tc5 timeout 300 python tools/bringup_vit_tc5.py run 2>&1 | grep -o '"time_24img.*' | cut -c1-300
echo "=== skinny ldg w8 m1"; VCB_SKINNY_WARPS=8 VCB_SKINNY_MODE=1 timeout 300 python tools/bench_skinny.py run 2>&1 | tail -1
echo "=== skinny ldg auto m1"; VCB_SKINNY_MODE=1 timeout 300 python tools/bench_skinny.py run 2>&1 | tail -1
for w in 2 4 8; do
echo "=== skinny tma warps=$w"; VCB_SKINNY_IMPL=tma VCB_SKINNY_TMA_WARPS=$w timeout 300 python tools/bench_skinny.py run 2>&1 | tail -1
done
echo "=== bench ldg"; timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench5_ldg.json 2> gpurun_out/bench5.err; python -c "
import json; d=json.load(open('gpurun_out/bench5_ldg.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['achieved']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench5.err
echo "=== bench tma"; VCB_SKINNY_IMPL=tma timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench5_tma.json 2> gpurun_out/bench5.err; python -c "
import json; d=json.load(open('gpurun_out/bench5_tma.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['achieved']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench5.err
