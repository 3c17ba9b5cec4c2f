#!/bin/bash
# executed FP64 instruction counts of the vector-model kernels at the BASELINE config shapes (under gpurun)
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
export BISIP_TIME_NOWARM=1
ncu --metrics $M --clock-control none -k regex:ensemble -c 1 --csv --log-file gpurun_out/r02_flops_dias.csv python tools/kernel_time.py --model dias --walkers 128 --spectra 1024 --steps 200 --reps 1
ncu --metrics $M --clock-control none -k regex:ensemble -c 1 --csv --log-file gpurun_out/r02_flops_shin.csv python tools/kernel_time.py --model shin --walkers 128 --spectra 1024 --steps 200 --reps 1
ncu --metrics $M --clock-control none -k regex:ensemble -c 1 --csv --log-file gpurun_out/r02_flops_colecole2_n20.csv python tools/kernel_time.py --model colecole --n-modes 2 --walkers 64 --n-freq 20 --spectra 1024 --steps 200 --reps 1
