#!/bin/bash
# round 2, call Q: ncu source-level sampling of the ViT attention kernel (where do the softmax warps sit?)
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'attn_vit_pair' -c 1 -o gpurun_out/prof_attn_vit_q -f python tools/ncu_ops.py attn > gpurun_out/ncu_attn_vit_q.log 2>&1; tail -1 gpurun_out/ncu_attn_vit_q.log
ncu -i gpurun_out/prof_attn_vit_q.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_attn_vit_q_source.csv 2>gpurun_out/ncu_src.err; wc -l gpurun_out/ncu_attn_vit_q_source.csv; head -3 gpurun_out/ncu_attn_vit_q_source.csv | cut -c1-600
rm -f gpurun_out/prof_attn_vit_q.ncu-rep
