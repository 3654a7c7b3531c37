#!/bin/bash
# round 2, call B: suite with the cluster-fused decode attention, microbenchmarks, decode timeline, ring-shape sweep
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/pytest_r2b.log 2>&1; tail -8 gpurun_out/pytest_r2b.log
echo "=== microbench"; timeout 120 python tools/microbench.py 2>&1 | tee gpurun_out/microbench_r2b.jsonl
echo "=== trace (cluster attention)"; timeout 300 python tools/trace_decode.py --label cluster --out gpurun_out/trace_decode_cluster.csv 2>&1 | tail -2
echo "=== trace (two-kernel attention)"; VCB_ATTN_DEC_CLUSTER=0 timeout 300 python tools/trace_decode.py --label split_combine --out gpurun_out/trace_decode_2k.csv 2>&1 | tail -2
echo "=== sweep"; timeout 600 python tools/sweep_decode.py 2>&1 | tee gpurun_out/sweep_decode_r2b.jsonl | tail -30
echo "=== done"
