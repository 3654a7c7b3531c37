#!/bin/bash
# round 2, call T: persistent ViT attention (item loop, running barrier phases, cross-item prefetch)
mkdir -p gpurun_out
echo "=== attention / tower tests (persistent)"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "attention_vit or vit_l_tower or tiny_against or tiny_quirks" 2>&1 | tail -5
echo "=== attention tests (VCB_VIT_PERSIST=0)"; VCB_VIT_PERSIST=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "attention_vit or vit_l_tower" 2>&1 | tail -3
echo "=== vit timing"; for pz in 1 0 1; do echo "persist=$pz"; VCB_VIT_PERSIST=$pz VCB_VIT_ATTN=pair timeout 200 python tools/bringup_vit_tc5.py run 2>&1 | tail -1 | python -c "import sys,json; l=sys.stdin.read(); d=json.loads(l[l.index('{'):]); print(d['time_24img'])"; done
echo "=== trace"; timeout 120 python tools/trace_vit_pair.py 2>&1 > gpurun_out/vit_pair_trace_t.txt; grep -v "^    [5-9] " gpurun_out/vit_pair_trace_t.txt | head -34
