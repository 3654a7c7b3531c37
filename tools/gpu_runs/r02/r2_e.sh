#!/bin/bash
# round 2, call E: ViT pair kernel with the MUFU turnstile (parity + trace + timing), decode attention occupancy / split variants
mkdir -p gpurun_out
echo "=== vit attention op tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -p no:cacheprovider -k "attention_vit or vit_l_tower or tiny_against" 2>&1 | tail -3
echo "=== vit pair trace (turnstile)"; timeout 120 python tools/trace_vit_pair.py 2>&1 | tail -14
echo "=== vit timing turnstile on/off"
for ts in 1 0; do VCB_VIT_TURNSTILE=$ts VCB_VIT_ATTN=pair timeout 120 python tools/bringup_vit_tc5.py run 2>&1 | tail -1 | python -c "import sys,json; l=sys.stdin.read(); d=json.loads(l[l.index('{'):]); print('turnstile=$ts', d['time_24img'])"; done
for cfg in "" "VCB_ATTN_DEC_MINB=4" "VCB_ATTN_DEC_CHUNK=128" "VCB_ATTN_DEC_CHUNK=128 VCB_ATTN_DEC_MINB=4" "VCB_ATTN_DEC_CHUNK=512 VCB_ATTN_DEC_MINB=4"; do
  echo "=== phases [$cfg]"; env $cfg timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1
done
echo "=== done"
