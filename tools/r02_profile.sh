#!/bin/bash
# Evidence run on one B200 (under gpurun): FP64 op counts of the vector-model kernels, DRAM traffic of the dominant kernel
# at bench size, the launch list of the bench command, one full-set capture of the default and the collapsed kernel.
set -x
mkdir -p gpurun_out
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:ensemble -c 1 --csv --log-file gpurun_out/r02_flops_dias.csv python tools/kernel_time.py --model dias --walkers 128 --spectra 1024 --steps 200 --reps 1
ncu --metrics $M --clock-control none -k regex:ensemble -c 1 --csv --log-file gpurun_out/r02_flops_shin.csv python tools/kernel_time.py --model shin --walkers 128 --spectra 1024 --steps 200 --reps 1
ncu --metrics $M --clock-control none -k regex:ensemble -c 1 --csv --log-file gpurun_out/r02_flops_colecole2_n20.csv python tools/kernel_time.py --model colecole --n-modes 2 --walkers 64 --n-freq 20 --spectra 1024 --steps 200 --reps 1
# DRAM bytes of the default kernel at bench size (one launch = the 12,500-spectra shard, 2000 steps, kept chain 100 steps)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ensemble_kernel -c 1 --csv --log-file gpurun_out/r02_traffic_bench_size.csv python tools/kernel_time.py --model decomp --precision fp64 --spectra 12500 --steps 2000 --reps 1
# full-set captures
ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -c 1 -o gpurun_out/r02_ensemble_decomp python tools/kernel_time.py --model decomp --precision fp64 --spectra 296 --steps 200 --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ensemble -c 1 -o gpurun_out/r02_ensemble_dias python tools/kernel_time.py --model dias --walkers 128 --spectra 1184 --steps 100 --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:column_stats -c 1 -o gpurun_out/r02_column_stats python tools/stats_bench.py --spectra 2000 > /dev/null 2>&1
# launch list of the bench command (short: 2 steps)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -20
