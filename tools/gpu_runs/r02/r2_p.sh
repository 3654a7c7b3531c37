#!/bin/bash
# round 2, call P (diagnostic builds, compile-time: VCB_EXTRA_CFLAGS=-DVCB_VIT_DIAG=n into diag_libs/lib_n.so): what the
# ViT attention's time does and does not depend on.  bits: 1 no MUFU (p = the FFMA result), 4 no max, 8 no row sums.
# Results are wrong by construction; timing only.
mkdir -p gpurun_out
cp vcoder_b200/libvcoder_b200.so /tmp/lib_prod.so
: > gpurun_out/vit_diag.txt
for cfg in "0 1" "1 1" "9 1" "4 1" "13 1" "0 0" "13 0"; do
  set -- $cfg
  cp diag_libs/lib_$1.so vcoder_b200/libvcoder_b200.so
  echo "=== diag $1 turnstile $2" >> gpurun_out/vit_diag.txt
  VCB_VIT_TURNSTILE=$2 timeout 120 python tools/trace_vit_pair.py >> gpurun_out/vit_diag.txt 2>&1
  VCB_VIT_TURNSTILE=$2 VCB_VIT_ATTN=pair timeout 200 python tools/bringup_vit_tc5.py run 2>&1 | tail -1 | python -c "import sys,json; l=sys.stdin.read(); d=json.loads(l[l.index('{'):]); print(d['time_24img'])" >> gpurun_out/vit_diag.txt
done
cp /tmp/lib_prod.so vcoder_b200/libvcoder_b200.so
grep -A4 "=== diag\|us_per_image" gpurun_out/vit_diag.txt | grep "=== diag\|us_per_image\|cta cycles\|    [23] "
