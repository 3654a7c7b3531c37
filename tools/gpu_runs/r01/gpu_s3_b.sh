#!/bin/bash
# Session-3 call B: tile-major packed decode weight stream — parity (bit-identical to the row-major stream), op-level A/B,
# then the whole job with VCB_SKINNY_PACKED=1 and =0 on the same box.
mkdir -p gpurun_out
echo "=== pytest packed"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -p no:cacheprovider -k "packed or skinny" 2>&1 | tail -8
echo "=== op-level A/B"; timeout 200 python tools/bench_skinny_packed.py 2>&1 | tail -3 | tee gpurun_out/skinny_packed_ab.json
for depth in 2 3; do
echo "=== op-level A/B depth=$depth"; VCB_SKINNY_TMA_DEPTH=$depth timeout 200 python tools/bench_skinny_packed.py 2>&1 | tail -1 | tee gpurun_out/skinny_packed_ab_d$depth.json
done
for pk in 1 0; do
echo "=== bench VCB_SKINNY_PACKED=$pk"; VCB_SKINNY_PACKED=$pk timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s3b_packed$pk.json 2> gpurun_out/bench_s3b_packed$pk.err; python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_s3b_packed$pk.json'))
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); print(d['roofline']['achieved'], d['roofline']['frac'], d['phases'])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_s3b_packed$pk.err
done
echo "=== done"
