#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== vit tc5"; VCB_VIT_ATTN=tc5 timeout 300 python tools/bringup_vit_tc5.py run 2>&1 | cut -c1-1500
for pf in 0 1; do
echo "=== skinny prefetch=$pf"; VCB_SKINNY_PREFETCH=$pf timeout 300 python tests/manual_bringup_engine.py ops 2>&1 | tr ',' '\n' | grep -E "skinny|GBs" | paste - - | cut -c1-200
done
for pf in 1 0; do
echo "=== bench prefetch=$pf"; VCB_SKINNY_PREFETCH=$pf timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench3_pf$pf.json 2> gpurun_out/bench3.err; python -c "
import json; d=json.load(open('gpurun_out/bench3_pf$pf.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['achieved']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench3.err
done
