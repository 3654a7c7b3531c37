#!/bin/bash
# round 2, call J: balanced page-streaming decode attention: op parity, suite, A/B vs the split kernel, timeline
mkdir -p gpurun_out
echo "=== decode attention op tests"; timeout 300 python -m pytest tests/test_gpu_longctx.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "decode_op" 2>&1 | tail -6
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/pytest_r2j.log 2>&1; tail -6 gpurun_out/pytest_r2j.log
for cfg in "" "VCB_ATTN_DEC=split"; do
  echo "=== phases [$cfg]"; env $cfg timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1
done
echo "=== trace"; timeout 300 python tools/trace_decode.py --label stream_attn --out gpurun_out/trace_decode_stream_attn.csv 2>&1 | tail -1
echo "=== 13B slice phases"; timeout 300 python bench.py --preset 13b --batch 4 --new-tokens 256 --steps 2 --warmup 2 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','phases','roofline')})" | cut -c1-700
echo "=== done"
