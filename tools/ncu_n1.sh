#!/bin/bash
# ncu evidence for the shipped one-GPU configuration (python bench.py defaults: 110 000 test users, head rows of songs with >= 150 listeners):
#   r02_launches.csv        every launch of one warm step with its device time (cold-cache, serialised: compare shares)
#   r02_traffic_raw.csv     dram bytes / L2 hit rate / time of the first ~700 scoring launches and a sample of the precompute launches
#   r02_head_rowsum.ncu-rep --set full of one UBM and one IBM head pass
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-k1-probe"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__inst_executed.sum.pct_of_peak_sustained_elapsed,launch__grid_size
$CMD > gpurun_out/ncu_n1_plain.json 2> gpurun_out/ncu_n1_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_n1_a.log 2>&1
ncu --metrics $M --clock-control none -k regex:'head_rowsum|tail_scatter|topk_kernel|mask_listened|head_fixup|zero_rows' -c 700 --csv --log-file gpurun_out/r02_traffic_scoring.csv $CMD > gpurun_out/ncu_n1_b.log 2>&1
ncu --metrics $M --clock-control none -k regex:'gram_head|pack_head' -c 2500 --csv --log-file gpurun_out/r02_traffic_precompute.csv $CMD > gpurun_out/ncu_n1_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_rowsum -s 1 -c 1 -f -o gpurun_out/r02_head_rowsum_ubm $CMD > gpurun_out/ncu_n1_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:head_rowsum -s 8 -c 1 -f -o gpurun_out/r02_head_rowsum_ibm $CMD > gpurun_out/ncu_n1_e.log 2>&1
ls -la gpurun_out/r02_*; tail -2 gpurun_out/ncu_n1_e.log
