#!/bin/bash
# Session-3 call H: last sanity pass at the final HEAD of the session — full GPU suite, smoke, decode/prefill phases.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 200 python -m pytest tests -m gpu -q --timeout 150 -p no:cacheprovider > gpurun_out/pytest_s3h.log 2>&1; tail -4 gpurun_out/pytest_s3h.log
echo "=== smoke"; timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 60 python tools/bench_decode.py 2>&1 | tail -1 | tee gpurun_out/phases_s3h.json
echo "=== done"
