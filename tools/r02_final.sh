#!/bin/bash
# Round-2 end-of-round verification on one B200 (under gpurun): evidence first (FP64 instruction counts, DRAM traffic at
# bench size, full-set capture, launch list -> profiles/ncu_summary.json refreshed ON THE BOX so that the bench line that
# follows quotes figures measured on the same build), then the GPU suite, smoke and both bench arms.
cd "$(dirname "$0")/.."
R=${ROUND:-r02c}; O=gpurun_out/$R; mkdir -p $O
timeout 400 bash tools/r02_flops.sh > $O/flops.log 2>&1
timeout 600 bash tools/r02_final_profile.sh > $O/final_profile.log 2>&1
python tools/r02_update_summary.py > $O/update_summary.log 2>&1; tail -8 $O/update_summary.log
cp profiles/ncu_summary.json $O/ncu_summary.json
cp gpurun_out/r02_flops_*.csv gpurun_out/r02_traffic_bench_size.csv gpurun_out/r02_launches.csv $O/ 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_ensemble_decomp.ncu-rep --name ensemble_decomp --spectra 296 --scale-spectra 12500 --note "final build: default FP64 two-stage DMMA under the warp-private sampler, 296 spectra, W=256 T=200 N=64 S=64" --out $O/${R}_ensemble_decomp > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ensemble -c 1 -f -o gpurun_out/r02_ensemble_dias python tools/kernel_time.py --model dias --walkers 128 --spectra 888 --steps 100 --reps 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02_ensemble_dias.ncu-rep --name ensemble_dias --spectra 888 --scale-spectra 1024 --note "final build: Dias, block-synchronous sampler, 128-thread CTAs x 6/SM, 4 frequencies in flight, 888 spectra, W=128 T=100 N=64" --out $O/${R}_ensemble_dias > /dev/null 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/${R}_pytest_gpu.log 2>&1; tail -3 $O/${R}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.log 2>&1; tail -1 $O/${R}_smoke.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $O/${R}_bench_reference.json 2> $O/bench_ref.err; cut -c1-300 $O/${R}_bench_reference.json
timeout 1200 python bench.py > $O/${R}_bench_n1.json 2> $O/bench.err; python tools/bench_digest.py $O/${R}_bench_n1.json
rm -f gpurun_out/*.ncu-rep.tmp; du -sh gpurun_out
