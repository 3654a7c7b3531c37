#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu full captures. Outputs under gpurun_out/.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench_ref.json
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
echo "=== ncu full: skinny"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -s 3 -c 12 -o gpurun_out/prof_skinny -f python tools/ncu_ops.py skinny > gpurun_out/ncu_skinny.log 2>&1; tail -2 gpurun_out/ncu_skinny.log
echo "=== ncu full: attn_vit"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_prefill -s 1 -c 2 -o gpurun_out/prof_attn_vit -f python tools/ncu_ops.py attn > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
echo "=== ncu full: gemm"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 0 -c 12 -o gpurun_out/prof_gemm -f python tools/ncu_ops.py gemm > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
ls -la gpurun_out
