#!/bin/bash
# round 2, final evidence at HEAD: suite, smoke, both bench arms, configs[1], ncu launch list + full captures, ViT trace, sweep
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/pytest_final.log 2>&1; tail -4 gpurun_out/pytest_final.log
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
echo "=== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_final.json'))
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
    print(d['roofline']['achieved'], d['roofline']['frac'], d['phases']); print(d['vit']); print(d['cpu_baseline']['value'], d['library_baseline'])
    print(d.get('ids_checksum'), d.get('graph_equals_eager'), d.get('e2e_equals_resident'))
    [print(k, v) for k, v in d['kernel_breakdown'].items()]
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_final.err
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_final.json 2> gpurun_out/ref_final.err; cut -c1-300 gpurun_out/ref_final.json
echo "=== configs[1]: VCoder-7B, B=1, 32 new tokens"; timeout 400 python bench.py --variant vcoder --batch 1 --new-tokens 32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vcoder_c1.json 2> gpurun_out/bench_vcoder_c1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_vcoder_c1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['library_baseline'], d['phases'])"; tail -2 gpurun_out/bench_vcoder_c1.err
echo "=== ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc5|skinny|attn_|rmsnorm|layernorm|rope_kv|im2col|argmax|sample|gather|cls_rows|advance|compact|depth_sum|fingerprint' -c 1700 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-library-baseline --new-tokens 8 > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/launches_final.csv
python tools/summarize_launches.py gpurun_out/launches_final.csv > gpurun_out/launches_summary_final.csv 2>&1; head -24 gpurun_out/launches_summary_final.csv
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size"
echo "=== ncu full: weight-streaming GEMM"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'skinny_stream' -c 15 -o gpurun_out/prof_skinny_final -f python tools/ncu_ops.py skinny > gpurun_out/ncu_skinny.log 2>&1; tail -1 gpurun_out/ncu_skinny.log
ncu -i gpurun_out/prof_skinny_final.ncu-rep --page raw --csv --metrics $M > gpurun_out/ncu_skinny_final.csv 2>/dev/null; cut -c1-400 gpurun_out/ncu_skinny_final.csv | tail -16
echo "=== ncu full: LLaMA attention (prefill tc5, decode stream)"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'attn_prefill_tc5|attn_decode_stream' -c 9 -o gpurun_out/prof_attn_llama_final -f python tools/ncu_ops.py attn_llama > gpurun_out/ncu_attn_llama.log 2>&1; tail -1 gpurun_out/ncu_attn_llama.log
ncu -i gpurun_out/prof_attn_llama_final.ncu-rep --page raw --csv --metrics $M > gpurun_out/ncu_attn_llama_final.csv 2>/dev/null; cut -c1-400 gpurun_out/ncu_attn_llama_final.csv | tail -10
echo "=== ncu full: ViT attention"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:'attn_vit_pair' -c 2 -o gpurun_out/prof_attn_vit_final -f python tools/ncu_ops.py attn > gpurun_out/ncu_attn_vit.log 2>&1; tail -1 gpurun_out/ncu_attn_vit.log
ncu -i gpurun_out/prof_attn_vit_final.ncu-rep --page raw --csv --metrics $M > gpurun_out/ncu_attn_vit_final.csv 2>/dev/null; cut -c1-400 gpurun_out/ncu_attn_vit_final.csv | tail -3
rm -f gpurun_out/prof_*_final.ncu-rep
echo "=== vit pair trace"; timeout 120 python tools/trace_vit_pair.py > gpurun_out/vit_pair_trace_final.txt 2>&1; head -4 gpurun_out/vit_pair_trace_final.txt
echo "=== tower sweep"; timeout 300 python tools/bench_tower_sweep.py --batches 1,8,32 --reps 3 2>&1 | tee gpurun_out/tower_sweep_final.jsonl | cut -c1-250
echo "=== done"
