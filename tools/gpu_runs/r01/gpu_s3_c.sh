#!/bin/bash
# Session-3 call C: L2 prefetch before the grid-dependency wait (skinny GEMM: VCB_SKINNY_PF groups; decode attention:
# VCB_ATTN_DEC_PF) — parity smoke, op-level A/B, decode-phase A/B of the whole 7B job, one process per variant.
mkdir -p gpurun_out
echo "=== parity with the prefetches on"; VCB_SKINNY_PF=4 VCB_ATTN_DEC_PF=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -p no:cacheprovider -k "skinny or tiny or real_width or packed" 2>&1 | tail -4
for pf in 0 2 4 8; do
echo "=== op-level VCB_SKINNY_PF=$pf"; VCB_SKINNY_PF=$pf timeout 120 python tools/bench_skinny_packed.py 2>&1 | tail -1 | tee gpurun_out/skinny_pf$pf.json
done
: > gpurun_out/decode_ab_s3c.jsonl
for cfg in "" "VCB_SKINNY_PF=2" "VCB_SKINNY_PF=4" "VCB_SKINNY_PF=8" "VCB_ATTN_DEC_PF=1" "VCB_SKINNY_PF=4 VCB_ATTN_DEC_PF=1" "VCB_SKINNY_PF=4 VCB_SKINNY_PACKED=1" ""; do
echo "=== decode [$cfg]"; env $cfg timeout 150 python tools/bench_decode.py 2>&1 | tail -1 | tee -a gpurun_out/decode_ab_s3c.jsonl
done
echo "=== done"
