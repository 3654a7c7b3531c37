#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu (tc5 default)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== vit tc5 v3b"; timeout 300 python tools/bringup_vit_tc5.py run 2>&1 | grep -o '"2x577x16.*' | cut -c1-800
echo "=== vit trace v3b"; timeout 120 python tools/trace_vit_tc5.py 2>&1 | tail -6
echo "=== vit pair (v4)"; VCB_VIT_ATTN=pair timeout 300 python tools/bringup_vit_tc5.py run 2>&1 | tail -3 | cut -c1-1200
echo "=== pytest attention (pair)"; VCB_VIT_ATTN=pair timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "attention or tiny" 2>&1 | tail -4
echo "=== bench"; timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err; python -c "
import json; d=json.load(open('gpurun_out/bench9.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench9.err
