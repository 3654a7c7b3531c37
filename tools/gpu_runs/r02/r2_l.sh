#!/bin/bash
# round 2, call L: co-resident decode kernels (VCB_DECODE_COLOC=1: one <=112 KiB CTA per SM per kernel, the next kernel's
# ring fills while the current one streams) — parity suite under the switch, decode A/B, device timeline
mkdir -p gpurun_out
echo "=== pytest -m gpu (COLOC=1)"; VCB_DECODE_COLOC=1 timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x -rf > gpurun_out/pytest_l.log 2>&1; tail -5 gpurun_out/pytest_l.log
run() { echo "--- $*"; env "$@" timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1 | tee -a gpurun_out/decode_coloc_ab.txt; }
: > gpurun_out/decode_coloc_ab.txt
run VCB_DECODE_COLOC=0
run VCB_DECODE_COLOC=1
run VCB_DECODE_COLOC=1 VCB_SKINNY_SLOTS=0 VCB_SKINNY_STREAM_DEPTH=2
run VCB_DECODE_COLOC=0 VCB_SKINNY_SLOTS=1 VCB_SKINNY_STREAM_DEPTH=3
run VCB_DECODE_COLOC=0 VCB_SKINNY_SLOTS=2 VCB_SKINNY_STREAM_DEPTH=2
run VCB_DECODE_COLOC=1 VCB_SKINNY_SLOTS=1 VCB_SKINNY_STREAM_DEPTH=2
echo "=== trace COLOC=1"; VCB_DECODE_COLOC=1 timeout 300 python tools/trace_decode.py --out gpurun_out/trace_decode_coloc.csv --label coloc 2>&1 | tail -30
echo "=== done"
