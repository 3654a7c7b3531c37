#!/bin/bash
# round 2, call M: RMSNorm fused into the balanced weight-streaming GEMM (Gram-diagonal statistics on the tensor pipe,
# VCB_FUSE_NORM=1) and decode attention with one CTA per (sequence, head) (no combine kernel, VCB_ATTN_DEC_WHOLE=1)
mkdir -p gpurun_out
echo "=== pytest -m gpu (FUSE_NORM=1 WHOLE=1)"; VCB_FUSE_NORM=1 VCB_ATTN_DEC_WHOLE=1 timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/pytest_m.log 2>&1; tail -8 gpurun_out/pytest_m.log
run() { echo "--- $*"; env "$@" timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1 | tee -a gpurun_out/decode_norm_whole_ab.txt; }
: > gpurun_out/decode_norm_whole_ab.txt
run VCB_FUSE_NORM=0
run VCB_FUSE_NORM=1
run VCB_ATTN_DEC_WHOLE=1
run VCB_ATTN_DEC_WHOLE=1 VCB_ATTN_DEC_WHOLE_RING=3
run VCB_FUSE_NORM=1 VCB_ATTN_DEC_WHOLE=1
run VCB_FUSE_NORM=1 VCB_ATTN_DEC_WHOLE=1 VCB_ATTN_DEC_WHOLE_RING=3
echo "=== trace both"; VCB_FUSE_NORM=1 VCB_ATTN_DEC_WHOLE=1 timeout 300 python tools/trace_decode.py --out gpurun_out/trace_decode_norm_whole.csv --label norm_whole 2>&1 | tail -3
echo "=== done"
