#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu (fused rope)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for f in 1 0; do
echo "=== bench fuse_rope=$f"; VCB_FUSE_ROPE=$f timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench11_$f.json 2> gpurun_out/bench11.err; python -c "
import json; d=json.load(open('gpurun_out/bench11_$f.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench11.err
done
