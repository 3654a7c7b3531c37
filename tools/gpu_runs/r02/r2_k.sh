#!/bin/bash
# round 2, run K: full GPU suite at HEAD (new: NF4, ragged shim, forward()/prepare_inputs seams), smoke, a short bench
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/pytest_k.log 2>&1; tail -25 gpurun_out/pytest_k.log
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
echo "=== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-library-baseline > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_k.json'))
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
    print(d['roofline']['achieved'], d['roofline']['frac'], d['phases'])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_k.err
echo "=== done"
