#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== skinny tma fused-norm microbench"; timeout 300 python - <<'PY'
import sys; sys.path.insert(0, "tools")
import bench_skinny as b
b.SHAPES = [(12288, 4096, 0, 1), (22016, 4096, 1, 1), (32000, 4096, 0, 1)]
print("fused", b.run())
b.SHAPES = [(12288, 4096, 0, 0), (22016, 4096, 1, 0), (32000, 4096, 0, 0)]
print("plain", b.run())
PY
for f in 1 0; do
echo "=== bench fuse_norm=$f"; VCB_FUSE_NORM=$f timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench10_$f.json 2> gpurun_out/bench10.err; python -c "
import json; d=json.load(open('gpurun_out/bench10_$f.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench10.err
done
