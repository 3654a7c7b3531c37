#!/bin/bash
# round 2, call N: box check (clocks / power), forward-seam tests after the fixes, "held-back stage" experiment:
# the last ring stage (GEMM) / page (attention) is requested right AFTER the grid-dependency wait instead of before it,
# so that HBM has requests in flight while the X fragments make their L2 round trip
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,power.limit,temperature.gpu --format=csv
echo "=== forward seam tests"; timeout 600 python -m pytest tests/test_gpu_forward_seam.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -3
run() { echo "--- $*"; env "$@" timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1 | tee -a gpurun_out/decode_hold_ab.txt; }
: > gpurun_out/decode_hold_ab.txt
run VCB_SKINNY_HOLD=0
run VCB_SKINNY_HOLD=1
run VCB_SKINNY_HOLD=1 VCB_SKINNY_STREAM_DEPTH=3
run VCB_SKINNY_HOLD=0 VCB_SKINNY_STREAM_DEPTH=3
run VCB_ATTN_DEC_HOLD=1
run VCB_ATTN_DEC_HOLD=1 VCB_SKINNY_HOLD=1
run VCB_SKINNY_HOLD=0
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv
echo "=== done"
