#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu (defaults: tma skinny)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== pytest -m gpu (tc5 attention)"; VCB_VIT_ATTN=tc5 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
echo "=== vit tc5"; VCB_VIT_ATTN=tc5 timeout 300 python tools/bringup_vit_tc5.py run 2>&1 | grep -o '"2x577x16.*' | cut -c1-700
for cfg in "4 2 4" "4 2 2" "4 3 2" "4 4 2" "8 2 4"; do
set -- $cfg
echo "=== skinny tma warps=$1 depth=$2 group=$3"; VCB_SKINNY_TMA_WARPS=$1 VCB_SKINNY_TMA_DEPTH=$2 VCB_SKINNY_TMA_GROUP=$3 timeout 300 python tools/bench_skinny.py run 2>&1 | tail -1
done
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err; python -c "
import json; d=json.load(open('gpurun_out/bench7.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['achieved'], d['cpu_baseline']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench7.err
echo "=== bench tc5 attn"; VCB_VIT_ATTN=tc5 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench7_tc5.json 2> gpurun_out/bench7.err; python -c "
import json; d=json.load(open('gpurun_out/bench7_tc5.json')); print(d['value'], d['ms_per_step']); print(d['kernel_breakdown']['attn_vit'])"
