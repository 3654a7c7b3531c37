#!/bin/bash
# Session-2 call A: parity + f1 tests at HEAD, smoke, bench (both arms), ncu launch list of one job.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_s2a.json 2> gpurun_out/bench_s2a.err; python -c "
import json; d=json.load(open('gpurun_out/bench_s2a.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); print(d['roofline']); print(d['phases']); print(d['vit']); print(d['cpu_baseline']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench_s2a.err
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_s2a.json 2> gpurun_out/bench_ref_s2a.err; cut -c1-600 gpurun_out/bench_ref_s2a.json
echo "=== ncu launch list (one job, 16 new tokens, own kernels)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc5|skinny|attn_|rmsnorm|layernorm|rope_kv|im2col|argmax|sample|gather|cls_rows|advance|compact|depth_sum' -c 6000 --csv --log-file gpurun_out/launches_s2a.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --new-tokens 16 > gpurun_out/bench_under_ncu_s2a.log 2>&1; tail -1 gpurun_out/bench_under_ncu_s2a.log | cut -c1-200
echo "=== ncu full: sample kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sample_kernel -c 2 -o gpurun_out/prof_sample -f python -m pytest tests/test_gpu_sampling.py -q -k "distribution" > gpurun_out/ncu_sample.log 2>&1; tail -2 gpurun_out/ncu_sample.log
