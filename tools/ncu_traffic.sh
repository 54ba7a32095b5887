#!/bin/bash
# DRAM traffic of the scoring kernels of ONE batch of the shipped one-GPU configuration (15 715 test users, 32 289 head rows), with
# --replay-mode application: every metric pass reruns the whole program, so nothing has to save / restore the ~100 GB the kernels write
# (kernel replay did not finish in 40 minutes on this configuration).  Bounded: own timeouts well inside the gpurun limit.
CMD="python bench.py --users 15715 --head-min-deg 150 --steps 1 --warmup 0 --no-cpu-baseline --no-k1-probe"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
timeout 300 $CMD > gpurun_out/ncu_traffic_plain.json 2> gpurun_out/ncu_traffic_plain.err || { echo plain run failed; exit 1; }
timeout -s INT 900 ncu --replay-mode application --metrics $M --clock-control none -k regex:'head_rowsum|tail_scatter|topk_kernel' -c 112 \
  --csv --log-file gpurun_out/r02_traffic_scoring.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r02_traffic_scoring.csv; tail -2 gpurun_out/ncu_traffic.log | cut -c1-300
