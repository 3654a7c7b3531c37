#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest skinny (tc5)"; VCB_SKINNY_IMPL=tc5 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "skinny" 2>&1 | tail -8
echo "=== pytest all (tc5)"; VCB_SKINNY_IMPL=tc5 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== skinny microbench tma"; timeout 300 python tools/bench_skinny.py run 2>&1 | tail -1
echo "=== skinny microbench tc5"; VCB_SKINNY_IMPL=tc5 timeout 300 python tools/bench_skinny.py run 2>&1 | tail -1
for ks in 1 2 4 8; do
echo "=== skinny microbench tc5 ksplit=$ks"; VCB_SKINNY_IMPL=tc5 VCB_SKINNY_TC5_KSPLIT=$ks timeout 300 python tools/bench_skinny.py run 2>&1 | tail -1
done
echo "=== bench tc5"; VCB_SKINNY_IMPL=tc5 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench13.json 2> gpurun_out/bench13.err; python -c "
import json; d=json.load(open('gpurun_out/bench13.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phases']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench13.err
