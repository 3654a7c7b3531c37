#!/bin/bash
# round 2, call H: MMA/TMA warps at the top warp ids in both tcgen05 attention kernels
mkdir -p gpurun_out
echo "=== attention tests"; timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_longctx.py -m gpu -q --timeout 200 -p no:cacheprovider -k "attention or vit_l_tower or tiny_against" 2>&1 | tail -4
echo "=== vit timing"; VCB_VIT_ATTN=pair timeout 200 python tools/bringup_vit_tc5.py run 2>&1 | tail -1 | python -c "import sys,json; l=sys.stdin.read(); d=json.loads(l[l.index('{'):]); print(d['time_24img'])"
echo "=== vit pair trace"; timeout 120 python tools/trace_vit_pair.py 2>&1 | tail -25
echo "=== phases"; timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1
echo "=== done"
