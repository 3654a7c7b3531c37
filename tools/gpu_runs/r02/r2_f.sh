#!/bin/bash
# round 2, call F: balanced multi-tile weight-streaming kernel (skinny_stream.cu): parity, A/B against the one-tile-per-CTA kernel
mkdir -p gpurun_out
echo "=== skinny + engine tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_longctx.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -5
for cfg in "" "VCB_SKINNY_IMPL=tma" "VCB_SKINNY_STREAM_DEPTH=3" "VCB_SKINNY_TPC=1"; do
  echo "=== phases [$cfg]"; env $cfg timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1
done
echo "=== trace"; timeout 300 python tools/trace_decode.py --label stream --out gpurun_out/trace_decode_stream.csv 2>&1 | tail -1
echo "=== done"
