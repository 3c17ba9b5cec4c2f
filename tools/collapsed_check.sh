#!/bin/bash
# developer check of the collapsed decomposition kernel: parity tests, timings at four shapes, one ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out/${ROUND:-r01h}; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_collapsed.py tests/test_gpu_fullsize.py -x -q -k collapsed 2>&1 | tail -3
{
for s in 64 256; do python tools/kernel_time.py --model decomp --precision fp64-collapsed --B 3552 --W 256 --T 300 --S $s --reps 3; done
python tools/kernel_time.py --model decomp --precision fp64-collapsed --B 3552 --W 128 --T 300 --reps 2
python tools/kernel_time.py --model decomp --precision fp64-collapsed --B 3552 --W 32 --T 500 --N 20 --S 40 --reps 2
} 2>&1 | tee $O/collapsed_times.log
ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -c 1 -f -o $O/collapsed3 python tools/kernel_time.py --model decomp --precision fp64-collapsed --B 444 --W 256 --T 200 --reps 0 > $O/ncu_collapsed3.log 2>&1
python tools/ncu_summary.py $O/collapsed3.ncu-rep --name ensemble_decomp_collapsed --spectra 444 --scale-spectra 12500 --note "round 1h, 444 spectra (3 CTAs/SM), W=256 T=200 N=64 S=64" --out $O/r01h_ensemble_decomp_collapsed > /dev/null
