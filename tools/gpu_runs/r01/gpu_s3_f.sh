#!/bin/bash
# Session-3 call F (final evidence at HEAD): full GPU suite, smoke, both bench arms, ncu launch list, ncu --set full of the
# CTA-pair GEMM + ViT attention, config-5 tower sweep, decode-attention prefetch A/B, config-4 per-GPU slice (13B, B=4, 256).
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 400 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_s3f.log 2>&1; tail -5 gpurun_out/pytest_s3f.log
echo "=== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench"; timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_s3f.json 2> gpurun_out/bench_s3f.err; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_s3f.json'))
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
    print(d['roofline']['achieved'], d['roofline']['frac'], d['phases']); print(d['vit']); print(d['cpu_baseline']['value'], d['cpu_baseline']['sample'][-40:])
    [print(k, v) for k, v in d['kernel_breakdown'].items()]
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_s3f.err
echo "=== bench reference arm"; timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_s3f.json 2> gpurun_out/bench_ref_s3f.err; cut -c1-200 gpurun_out/bench_ref_s3f.json
echo "=== ncu launch list"
timeout 220 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc5|skinny|attn_|rmsnorm|layernorm|rope_kv|im2col|argmax|sample|gather|cls_rows|advance|compact|depth_sum|fingerprint' -c 1500 --csv --log-file gpurun_out/launches_s3f.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --new-tokens 8 > gpurun_out/bench_under_ncu_s3f.log 2>&1; wc -l gpurun_out/launches_s3f.csv
echo "=== ncu full: GEMM (pair + 1-CTA) and ViT attention"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc5' -c 8 -o gpurun_out/prof_gemm_s3f -f python tools/ncu_ops.py gemm > gpurun_out/ncu_gemm_s3f.log 2>&1; tail -1 gpurun_out/ncu_gemm_s3f.log
ncu -i gpurun_out/prof_gemm_s3f.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__registers_per_thread,launch__cluster_size > gpurun_out/ncu_gemm_s3f.csv 2>/dev/null; cut -c1-330 gpurun_out/ncu_gemm_s3f.csv | tail -9
timeout 60 ncu --set full --clock-control none -k regex:'attn_vit_pair' -c 2 -o gpurun_out/prof_attn_s3f -f python tools/ncu_ops.py attn > gpurun_out/ncu_attn_s3f.log 2>&1; tail -1 gpurun_out/ncu_attn_s3f.log
ncu -i gpurun_out/prof_attn_s3f.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active > gpurun_out/ncu_attn_s3f.csv 2>/dev/null; cut -c1-300 gpurun_out/ncu_attn_s3f.csv | tail -3
echo "=== tower sweep"; timeout 100 python tools/bench_tower_sweep.py > gpurun_out/tower_sweep_s3f.jsonl 2> gpurun_out/tower_sweep_s3f.err; cut -c1-260 gpurun_out/tower_sweep_s3f.jsonl
echo "=== decode attention prefetch A/B"
: > gpurun_out/decode_ab_s3f.jsonl
for cfg in "VCB_ATTN_DEC_PF=0" "VCB_ATTN_DEC_PF=1" "VCB_ATTN_DEC_PF=0" "VCB_ATTN_DEC_PF=1"; do env $cfg timeout 100 python tools/bench_decode.py 2>&1 | tail -1 | tee -a gpurun_out/decode_ab_s3f.jsonl; done
echo "=== config-4 slice: 13B, batch 4, 256 new tokens"
timeout 240 python bench.py --preset 13b --batch 4 --new-tokens 256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_13b_s3f.json 2> gpurun_out/bench_13b_s3f.err; cut -c1-260 gpurun_out/bench_13b_s3f.json; tail -2 gpurun_out/bench_13b_s3f.err
echo "=== done"
