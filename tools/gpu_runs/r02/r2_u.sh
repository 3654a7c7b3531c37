#!/bin/bash
# round 2, call U: RoPE + KV append fused into the QKV GEMM epilogue at prefill (kActRopeKv)
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x -rf > gpurun_out/pytest_u.log 2>&1; tail -4 gpurun_out/pytest_u.log
run() { echo "--- $*"; env "$@" timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1 | tee -a gpurun_out/prefill_rope_ab.txt; }
: > gpurun_out/prefill_rope_ab.txt
run VCB_PREFILL_ROPE=kernel
run VCB_PREFILL_ROPE=fused
run VCB_PREFILL_ROPE=kernel
run VCB_PREFILL_ROPE=fused
