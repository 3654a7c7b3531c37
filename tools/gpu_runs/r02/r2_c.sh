#!/bin/bash
# round 2, call C: bring-up of the tcgen05 prefill attention and the bulk-copy decode attention (op tests first, under a
# short timeout), then the suite, decode timelines (new attention; chained norm) and the phase timing
mkdir -p gpurun_out
echo "=== op tests"; timeout 300 python -m pytest tests/test_gpu_longctx.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "op" > gpurun_out/pytest_r2c_ops.log 2>&1; tail -15 gpurun_out/pytest_r2c_ops.log
echo "=== legacy prefill attention op tests"; VCB_PREFILL_ATTN=legacy timeout 300 python -m pytest tests/test_gpu_longctx.py -m gpu -q --timeout 120 -p no:cacheprovider -k "prefill_op" 2>&1 | tail -3
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/pytest_r2c.log 2>&1; tail -8 gpurun_out/pytest_r2c.log
echo "=== trace (bulk-copy attention)"; timeout 300 python tools/trace_decode.py --label tma_attn --out gpurun_out/trace_decode_tma.csv 2>&1 | tail -1
echo "=== trace (chained norm)"; VCB_DECODE_NORM_CHAIN=1 timeout 300 python tools/trace_decode.py --label tma_attn_chain --out gpurun_out/trace_decode_chain.csv 2>&1 | tail -1
echo "=== phases"; timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1
echo "=== phases, legacy prefill attention + ldg decode attention"; VCB_PREFILL_ATTN=legacy VCB_ATTN_DEC=ldg timeout 300 python tools/bench_decode.py --reps 3 2>&1 | tail -1
echo "=== done"
