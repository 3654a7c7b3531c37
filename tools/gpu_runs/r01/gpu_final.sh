#!/bin/bash
# Round-end evidence run: parity tests, smoke, bench (both arms), ncu launch list (own kernels only) and full captures.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== vit attention timing"; timeout 300 python tools/bringup_vit_tc5.py run 2>&1 | grep -o '"time_24img.*' | cut -c1-300
echo "=== vit trace"; timeout 120 python tools/trace_vit_tc5.py 2>&1 | tail -6
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-900 gpurun_out/bench_ref.json
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); print(d['roofline']); print(d['cpu_baseline']['value']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"; tail -3 gpurun_out/bench.err
echo "=== ncu launch list (one full step, own kernels)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc5|skinny|attn_|rmsnorm|layernorm|rope_kv|im2col|argmax|gather|cls_rows|advance|compact|depth_sum' -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --new-tokens 16 > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
echo "=== ncu full: skinny tma"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skinny -s 5 -c 10 -o gpurun_out/prof_skinny -f python tools/ncu_ops.py skinny > gpurun_out/ncu_skinny.log 2>&1; tail -1 gpurun_out/ncu_skinny.log
echo "=== ncu full: attn tc5"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_vit_tc5 -s 1 -c 1 -o gpurun_out/prof_attn_vit_tc5 -f python tools/ncu_ops.py attn > gpurun_out/ncu_attn_tc5.log 2>&1; tail -1 gpurun_out/ncu_attn_tc5.log
echo "=== ncu full: gemm"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 0 -c 12 -o gpurun_out/prof_gemm -f python tools/ncu_ops.py gemm > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log
