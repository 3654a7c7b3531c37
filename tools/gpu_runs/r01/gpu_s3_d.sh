#!/bin/bash
# Session-3 call D: bring-up of the CTA-pair (cta_group::2) GEMM: correctness per case in its own process (block_n = 512
# selects the pair kernel), then timing against the 1-CTA 128x256 kernel on the prefill / tower shapes.
mkdir -p gpurun_out
: > gpurun_out/gemm_pair_s3d.log
run() { echo "case $*" | tee -a gpurun_out/gemm_pair_s3d.log; timeout 60 python tools/bringup_gemm.py "$@" 2>&1 | tail -1 | cut -c1-600 | tee -a gpurun_out/gemm_pair_s3d.log; }
# M N K act bias res bn time
run 256 256 64 0 0 0 512 0
run 256 256 256 0 0 0 512 0
run 512 512 640 0 1 0 512 0
run 300 512 640 1 1 1 512 0
run 1000 1024 1024 2 1 0 512 0
run 1000 1024 1024 3 0 0 512 0
run 2414 4096 4096 0 0 1 512 0
for shape in "9656 12288 4096 0 0 0" "9656 4096 4096 0 0 1" "9656 22016 4096 3 0 0" "9656 4096 11008 0 0 1" "13848 3072 1024 0 1 0" "13848 4096 1024 1 1 0" "13848 1024 4096 0 1 1"; do
  run $shape 256 1
  run $shape 512 1
done
echo "=== done"
