#!/bin/bash
# Final evidence for the dominant kernel of the bench (default FP64 two-stage DMMA under the warp-private sampler):
# DRAM bytes at bench size, one full-set capture, and the launch list of a short bench run.
set -x
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ensemble -c 1 --csv --log-file gpurun_out/r02_traffic_bench_size.csv python tools/kernel_time.py --model decomp --precision fp64 --spectra 12500 --steps 2000 --reps 1
ncu --set full --clock-control none --import-source on -k regex:ensemble -c 1 -o gpurun_out/r02_ensemble_decomp python tools/kernel_time.py --model decomp --precision fp64 --spectra 296 --steps 200 --reps 1 > /dev/null 2>&1
BISIP_BENCH_CONFIGS=0 BISIP_BENCH_STRONG_SPECTRA=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_under_ncu.log 2>&1
