#!/bin/bash
# round 2, call O: timing-only probe — RMSNorm statistics from the Gram diagonal inside the balanced GEMM with the norm
# weight assumed folded into W (no per-stage weight loads), rmsnorm launches skipped (VCB_FUSE_NORM=1)
mkdir -p gpurun_out
run() { echo "--- $*"; env "$@" timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1 | tee -a gpurun_out/decode_gram_probe.txt; }
: > gpurun_out/decode_gram_probe.txt
run VCB_FUSE_NORM=0
run VCB_FUSE_NORM=1
run VCB_FUSE_NORM=0
run VCB_FUSE_NORM=1
echo "=== trace"; VCB_FUSE_NORM=1 timeout 300 python tools/trace_decode.py --out gpurun_out/trace_decode_gram_probe.csv --label gram_probe 2>&1 | tail -2
echo "=== done"
