#!/bin/bash
# round 2, call G: ViT attention with half-block softmax (parity, phase trace, timing)
mkdir -p gpurun_out
echo "=== vit attention tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -p no:cacheprovider -k "attention_vit or vit_l_tower or tiny_against" 2>&1 | tail -5
echo "=== extra shapes"; VCB_VIT_ATTN=pair timeout 200 python tools/bringup_vit_tc5.py run 2>&1 | tail -1 | cut -c1-1200
echo "=== vit pair trace"; timeout 120 python tools/trace_vit_pair.py 2>&1 | tail -26
echo "=== done"
