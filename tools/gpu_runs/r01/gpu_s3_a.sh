#!/bin/bash
# Session-3 call A (one box, ~14 min): full GPU parity suite at HEAD (f3 prefix cache, f2/f4 re-check), smoke, both bench
# arms, ncu launch list of one job, config-5 tower sweep, ncu --set full of the two kernels the rooflines quote.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q --timeout 300 --durations=12 -p no:cacheprovider > gpurun_out/pytest_s3a.log 2>&1; tail -40 gpurun_out/pytest_s3a.log
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; timeout 420 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_s3a.json 2> gpurun_out/bench_s3a.err; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_s3a.json'))
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
    print(d['roofline']); print(d['phases']); print(d['vit']); print(d['cpu_baseline'])
    [print(k, v) for k, v in d['kernel_breakdown'].items()]
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_s3a.err
echo "=== bench reference arm"; timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_s3a.json 2> gpurun_out/bench_ref_s3a.err; cut -c1-400 gpurun_out/bench_ref_s3a.json
echo "=== tower sweep (config 5)"; timeout 150 python tools/bench_tower_sweep.py > gpurun_out/tower_sweep_s3a.jsonl 2> gpurun_out/tower_sweep_s3a.err; cat gpurun_out/tower_sweep_s3a.jsonl | cut -c1-330; tail -2 gpurun_out/tower_sweep_s3a.err
echo "=== ncu launch list (one job, 16 new tokens, own kernels)"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc5|skinny|attn_|rmsnorm|layernorm|rope_kv|im2col|argmax|sample|gather|cls_rows|advance|compact|depth_sum|fingerprint' -c 6000 --csv --log-file gpurun_out/launches_s3a.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --new-tokens 16 > gpurun_out/bench_under_ncu_s3a.log 2>&1; tail -1 gpurun_out/bench_under_ncu_s3a.log | cut -c1-200; wc -l gpurun_out/launches_s3a.csv
echo "=== ncu full: skinny (1 launch per shape) + ViT attention"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'skinny_tma_kernel|attn_vit_pair_kernel' -c 8 -o gpurun_out/prof_s3a -f python tools/ncu_ops.py > gpurun_out/ncu_full_s3a.log 2>&1; tail -2 gpurun_out/ncu_full_s3a.log
ncu -i gpurun_out/prof_s3a.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active > gpurun_out/ncu_full_s3a.csv 2>/dev/null; cut -c1-260 gpurun_out/ncu_full_s3a.csv | head -14
echo "=== done"
