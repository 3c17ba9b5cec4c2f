#!/bin/bash
# round 2c evidence run (one B200, under gpurun): DFMA + companion-instruction issue test, FP64 instruction counts of the
# vector-model kernels after the exp-table / branch-free-loop change, timings of the final vector kernels
cd "$(dirname "$0")/.."
O=gpurun_out/r02c; mkdir -p $O
tools/peaks --mix > $O/peaks_mix.json; cat $O/peaks_mix.json
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sampler.py tests/test_gpu_edges.py tests/test_gpu_cython_shim.py -m gpu -x -q 2>&1 | tail -3
timeout 300 bash tools/r02_flops.sh > $O/flops.log 2>&1; tail -2 $O/flops.log
for cfg in "--model dias --walkers 128 --spectra 2368" "--model shin --walkers 128 --spectra 1776" "--model colecole --n-modes 1 --walkers 128 --spectra 2368" "--model colecole --n-modes 2 --walkers 128 --spectra 1776" "--model colecole --n-modes 2 --walkers 64 --n-freq 20 --spectra 3552" "--model dias --walkers 32 --n-freq 20 --spectra 9472" "--model dias --walkers 128 --spectra 1024" "--model shin --walkers 128 --spectra 1024"; do
  timeout 120 python tools/kernel_time.py $cfg --steps 500 --reps 3 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["model"], "modes", j["n_modes"], "W", j["walkers"], "N", j["n_freq"], "B", j["spectra"], "%.3e" % j["evals_per_s"])'
done | tee $O/vec_final.log
