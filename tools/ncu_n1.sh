#!/bin/bash
# ncu launch list of the shipped one-GPU command (python bench.py defaults: 110 000 test users, head rows of songs with >= 150 listeners):
# every launch with its device time (cold-cache, serialised: compare shares).  ONLY the single-pass time metric: kernel-replay passes with
# several metrics have to save / restore the ~100 GB these kernels write and did not finish in 40 minutes (the kill took the GPU down);
# DRAM traffic is captured by tools/ncu_traffic.sh with application replay on one batch instead.
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-k1-probe"
timeout 300 $CMD > gpurun_out/ncu_n1_plain.json 2> gpurun_out/ncu_n1_plain.err || exit 1
timeout -s INT 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_n1_a.log 2>&1
ls -la gpurun_out/r02_launches.csv
