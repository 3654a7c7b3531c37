#!/bin/bash
mkdir -p gpurun_out
for ts in 0 1; do
echo "=== vit timing turnstile=$ts"; VCB_VIT_TURNSTILE=$ts VCB_VIT_ATTN=pair timeout 200 python tools/bringup_vit_tc5.py run 2>&1 | tail -1 | python -c "import sys,json; l=sys.stdin.read(); d=json.loads(l[l.index('{'):]); print({k:(v if k=='time_24img' else v['rel_rms']) for k,v in d.items()})"
done
echo "=== vit pair trace turnstile"; VCB_VIT_TURNSTILE=1 timeout 120 python tools/trace_vit_pair.py 2>&1 | tail -25
echo "=== done"
