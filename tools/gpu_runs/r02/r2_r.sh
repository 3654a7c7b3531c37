#!/bin/bash
# round 2, call R: ViT attention — TMA requests in order of first use, no P-buffer barrier at even halves
mkdir -p gpurun_out
echo "=== attention / tower tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "attention_vit or vit_l_tower or tiny_against or tiny_quirks" 2>&1 | tail -3
echo "=== vit timing"; for i in 1 2; do VCB_VIT_ATTN=pair timeout 200 python tools/bringup_vit_tc5.py run 2>&1 | tail -1 | python -c "import sys,json; l=sys.stdin.read(); d=json.loads(l[l.index('{'):]); print(d['time_24img'])"; done
echo "=== trace"; timeout 120 python tools/trace_vit_pair.py 2>&1 | tee gpurun_out/vit_pair_trace_r.txt | head -14
