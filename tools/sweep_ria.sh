#!/bin/bash
# developer sweep (round 2d): bin slots of the key ranking inside the second accept phase (default) vs in their own phase
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02d
run() { echo "lib=${BISIP_B200_LIB:-default} $(timeout 120 python tools/kernel_time.py "$@" --steps 500 --reps 4 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["model"], j["precision"], "modes", j["n_modes"], "W", j["walkers"], "S", j["n_tau"], "B", j["spectra"], "%.3e" % j["evals_per_s"])')"; }
for rep in 1 2; do
for lib in "" bisip_b200/csrc/libbisip_b200_ria0.so; do
  if [ -n "$lib" ]; then export BISIP_B200_LIB=$lib; else unset BISIP_B200_LIB; fi
  run --model dias --walkers 128 --spectra 2368
  run --model dias --walkers 256 --spectra 1184
  run --model shin --walkers 128 --spectra 1776
  run --model colecole --n-modes 2 --walkers 128 --spectra 1776
  run --model decomp --precision 3xtf32 --spectra 592
  run --model decomp --precision tf32 --spectra 592
  run --model decomp --precision fp64 --n-tau 256 --spectra 296
done
done 2>&1 | tee gpurun_out/r02d/sweep_ria.log
