#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== skinny tma microbench"; timeout 300 python - <<'PY'
import sys; sys.path.insert(0, "tools")
import bench_skinny as b
print("plain", b.run())
b.SHAPES = [(12288, 4096, 0, 1), (22016, 4096, 1, 1), (32000, 4096, 0, 1)]
print("fused", b.run())
PY
for f in 0 1; do
echo "=== bench fuse_norm=$f"; VCB_FUSE_NORM=$f timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench12_$f.json 2> gpurun_out/bench12.err; python -c "
import json; d=json.load(open('gpurun_out/bench12_$f.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phases']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench12.err
done
