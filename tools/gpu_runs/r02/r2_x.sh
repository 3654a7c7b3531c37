#!/bin/bash
# round 2, call X: ncu on the tower-shaped tcgen05 GEMMs (K = 1024 / 4096 at M = 13848) next to the prefill shapes
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:'gemm_tc5' -c 12 -o gpurun_out/prof_gemm_x -f python tools/ncu_ops.py gemm > gpurun_out/ncu_gemm_x.log 2>&1; tail -1 gpurun_out/ncu_gemm_x.log
ncu -i gpurun_out/prof_gemm_x.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_x_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/ncu_gemm_x_raw.csv')))
hdr = rows[0]
keys = ['Kernel Name', 'launch__grid_size', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']
idx = [hdr.index(k) for k in keys if k in hdr]
print([hdr[i] for i in idx])
for r in rows[2:]:
    print([r[i][:60] for i in idx])
PY
rm -f gpurun_out/prof_gemm_x.ncu-rep
