#!/bin/bash
# Session-3 call G: chained RMSNorm in the decode step (no rmsnorm kernels) — full GPU suite, A/B of the decode phase,
# the headline bench at the new default, and a fresh launch list.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/pytest_s3g.log 2>&1; tail -6 gpurun_out/pytest_s3g.log
: > gpurun_out/decode_ab_s3g.jsonl
for cfg in "VCB_DECODE_NORM_CHAIN=0" "VCB_DECODE_NORM_CHAIN=1"; do env $cfg timeout 100 python tools/bench_decode.py 2>&1 | tail -1 | tee -a gpurun_out/decode_ab_s3g.jsonl; done
echo "=== bench"; timeout 240 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_s3g.json 2> gpurun_out/bench_s3g.err; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_s3g.json'))
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
    print(d['roofline']['achieved'], d['roofline']['frac'], d['phases']); print(d['cpu_baseline']['value'])
    [print(k, v) for k, v in d['kernel_breakdown'].items()]
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_s3g.err
echo "=== ncu launch list"
timeout 110 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc5|skinny|attn_|rmsnorm|layernorm|rope_kv|im2col|argmax|sample|gather|cls_rows|advance|compact|depth_sum|fingerprint' -c 1100 --csv --log-file gpurun_out/launches_s3g.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --new-tokens 6 > gpurun_out/bench_under_ncu_s3g.log 2>&1; wc -l gpurun_out/launches_s3g.csv
echo "=== done"
