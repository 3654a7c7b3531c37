#!/bin/bash
# Session-3 call E: grouped tile rasterisation in both GEMM kernels — parity, then the same timing cases as call D
# (row-major order: profiles/r01_gemm_pair_s3.log), then the whole job's prefill / decode phases.
mkdir -p gpurun_out
echo "=== pytest gemm + engine parity"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -p no:cacheprovider -k "gemm_tc5 or tiny or real_width or vit_l" 2>&1 | tail -4
: > gpurun_out/gemm_raster_s3e.log
run() { echo "case $*" | tee -a gpurun_out/gemm_raster_s3e.log; timeout 60 python tools/bringup_gemm.py "$@" 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/gemm_raster_s3e.log; }
for shape in "9656 12288 4096 0 0 0" "9656 4096 4096 0 0 1" "9656 22016 4096 3 0 0" "9656 4096 11008 0 0 1" "13848 3072 1024 0 1 0" "13848 4096 1024 1 1 0" "13848 1024 4096 0 1 1" "13848 1024 1024 0 1 1" "4616 4096 1024 2 1 0" "4616 4096 4096 0 1 0"; do
  run $shape 256 1
  run $shape 512 1
done
echo "=== job phases"; timeout 150 python tools/bench_decode.py 2>&1 | tail -1 | tee gpurun_out/phases_s3e.json
echo "=== done"
