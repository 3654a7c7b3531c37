#!/bin/bash
# round 2, call D: last-CTA merge in the decode attention, chained RMSNorm v2 (rstd finalised by the producer's last CTA),
# ViT pair-kernel phase trace, tower sweep to 256 samples
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/pytest_r2d.log 2>&1; tail -8 gpurun_out/pytest_r2d.log
for cfg in "" "VCB_ATTN_DEC=ldg2" "VCB_DECODE_NORM_CHAIN=1" "VCB_DECODE_NORM_CHAIN=1 VCB_ATTN_DEC=ldg2"; do
  echo "=== phases [$cfg]"; env $cfg timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1
done
echo "=== trace default"; timeout 300 python tools/trace_decode.py --label merge_last_cta --out gpurun_out/trace_decode_d.csv 2>&1 | tail -1
echo "=== trace chain v2"; VCB_DECODE_NORM_CHAIN=1 timeout 300 python tools/trace_decode.py --label chain_v2 --out gpurun_out/trace_decode_d_chain.csv 2>&1 | tail -1
echo "=== vit pair trace"; timeout 120 python tools/trace_vit_pair.py 2>&1 | tail -16
echo "=== tower sweep"; timeout 600 python tools/bench_tower_sweep.py --reps 3 2>&1 | tee gpurun_out/tower_sweep_r2d.jsonl | cut -c1-260
echo "=== done"
