#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== skinny sweep"; timeout 1200 python tools/bench_skinny.py 2>&1 | cut -c1-300
echo "=== bench default"; timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; python -c "
import json; d=json.load(open('gpurun_out/bench4.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['achieved']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench4.err
echo "=== ncu attn tc5"
VCB_VIT_ATTN=tc5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_vit_tc5 -s 1 -c 1 -o gpurun_out/prof_attn_vit_tc5 -f python tools/ncu_ops.py attn > gpurun_out/ncu_attn_tc5.log 2>&1; tail -2 gpurun_out/ncu_attn_tc5.log
