#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "=== vit tc5 bring-up"; timeout 900 python tools/bringup_vit_tc5.py 2>&1 | cut -c1-1200
echo "=== skinny mode 0"; VCB_SKINNY_MODE=0 timeout 300 python tests/manual_bringup_engine.py ops 2>&1 | tr ',' '\n' | grep -E "skinny|GBs" | paste - - | cut -c1-200
echo "=== skinny mode 1"; VCB_SKINNY_MODE=1 timeout 300 python tests/manual_bringup_engine.py ops 2>&1 | tr ',' '\n' | grep -E "skinny|GBs" | paste - - | cut -c1-200
echo "=== bench (PDL on)"; timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; python -c "
import json; d=json.load(open('gpurun_out/bench2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['achieved']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench2.err
echo "=== bench (PDL off)"; VCB_PDL=0 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench2_nopdl.json 2>> gpurun_out/bench2.err; python -c "
import json; d=json.load(open('gpurun_out/bench2_nopdl.json')); print(d['value'], d['ms_per_step'])"
