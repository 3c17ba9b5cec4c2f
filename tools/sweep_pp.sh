#!/bin/bash
# developer sweep (round 2e): pair-propose on / off in the block-synchronous sampler, same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02e
run() { echo "lib=${BISIP_B200_LIB:-default} $(timeout 120 python tools/kernel_time.py "$@" --steps 500 --reps 4 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["model"], j["precision"], "W", j["walkers"], "S", j["n_tau"], "B", j["spectra"], "%.3e" % j["evals_per_s"])')"; }
for rep in 1 2; do
for lib in "" bisip_b200/csrc/libbisip_b200_pp0.so; do
  if [ -n "$lib" ]; then export BISIP_B200_LIB=$lib; else unset BISIP_B200_LIB; fi
  run --model decomp --precision 3xtf32 --spectra 592
  run --model decomp --precision tf32 --spectra 592
  run --model decomp --precision fp64 --n-tau 256 --spectra 296
  run --model decomp --precision 3xtf32 --n-tau 256 --spectra 296
done
done 2>&1 | tee gpurun_out/r02e/sweep_pp.log
