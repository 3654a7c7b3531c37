#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== vit tc5"; timeout 300 python tools/bringup_vit_tc5.py run 2>&1 | grep -o '"2x577x16.*' | cut -c1-800
echo "=== vit legacy"; VCB_VIT_ATTN=legacy timeout 300 python tools/bringup_vit_tc5.py run 2>&1 | grep -o '"time_24img.*' | cut -c1-300
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench8.json 2> gpurun_out/bench8.err; python -c "
import json; d=json.load(open('gpurun_out/bench8.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['achieved'], d['cpu_baseline']['value']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench8.err
echo "=== ncu attn tc5"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_vit_tc5 -s 1 -c 1 -o gpurun_out/prof_attn_vit_tc5 -f python tools/ncu_ops.py attn > gpurun_out/ncu_attn_tc5.log 2>&1; tail -1 gpurun_out/ncu_attn_tc5.log
