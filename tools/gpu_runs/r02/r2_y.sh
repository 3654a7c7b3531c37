#!/bin/bash
# round 2, call Y: eight epilogue warps in the tcgen05 GEMMs — full suite, then A/B against the four-warp build
# (diag_libs/lib_old.so) on the tower (24 images) and on the prefill phase
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_y.log 2>&1; tail -3 gpurun_out/pytest_y.log
cp vcoder_b200/libvcoder_b200.so /tmp/lib_new.so
for v in old new old new; do
  if [ $v = old ]; then cp diag_libs/lib_old.so vcoder_b200/libvcoder_b200.so; else cp /tmp/lib_new.so vcoder_b200/libvcoder_b200.so; fi
  echo "--- $v"; timeout 300 python tools/bench_tower_sweep.py --batches 8 --reps 5 2>&1 | head -1 | cut -c1-200
  timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1
done | tee gpurun_out/gemm_epilogue8_ab.txt
cp /tmp/lib_new.so vcoder_b200/libvcoder_b200.so
