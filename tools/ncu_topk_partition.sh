#!/bin/bash
# ncu --set full of one UBM and one IBM top-k launch of an emulated song partition (1/8 of the songs, all test users)
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-k1-probe --emulate-song-partition 0/8 --head-min-deg 125"
timeout 240 $CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err &&
timeout -s INT 400 ncu --set full --clock-control none --import-source on -k regex:topk_kernel -s 5 -c 1 -f -o gpurun_out/r02b_topk_part_ubm $CMD > gpurun_out/ncu_a.log 2>&1
timeout -s INT 400 ncu --set full --clock-control none --import-source on -k regex:topk_kernel -s 29 -c 1 -f -o gpurun_out/r02b_topk_part_ibm $CMD > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_a.log
