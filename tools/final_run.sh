#!/bin/bash
# Round-end verification on one B200 (run through gpurun): GPU test suite, smoke, both bench arms, the ncu launch
# list of the bench command and one --set full capture of the collapsed decomposition kernel.
cd "$(dirname "$0")/.."
R=${ROUND:-r01i}; O=gpurun_out/$R; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/${R}_pytest_gpu.log 2>&1; tail -3 $O/${R}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.log 2>&1; tail -1 $O/${R}_smoke.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $O/${R}_bench_reference.json 2> $O/bench_ref.err
timeout 900 python bench.py > $O/${R}_bench_n1.json 2> $O/bench.err; cut -c1-400 $O/${R}_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -c 1 -f"
timeout 300 $NCU -o $O/collapsed python tools/kernel_time.py --model decomp --precision fp64-collapsed --B 444 --W 256 --T 200 --reps 0 > $O/ncu_collapsed.log 2>&1
python tools/ncu_summary.py $O/collapsed.ncu-rep --name ensemble_decomp_collapsed --spectra 444 --scale-spectra 12500 --note "final build, 444 spectra (3 CTAs/SM), W=256 T=200 N=64 S=64" --out $O/${R}_ensemble_decomp_collapsed > /dev/null
du -sh gpurun_out
