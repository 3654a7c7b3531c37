#!/bin/bash
# round 2, call S: ViT attention — nanosleep between the mbarrier probes of the issuer / producer threads
for ns in 0 32 100 300 1000 0; do
  echo "--- VCB_VIT_POLL_NS=$ns"; VCB_VIT_POLL_NS=$ns VCB_VIT_ATTN=pair timeout 200 python tools/bringup_vit_tc5.py run 2>&1 | tail -1 | python -c "import sys,json; l=sys.stdin.read(); d=json.loads(l[l.index('{'):]); print(d['time_24img'])"
done
