#!/bin/bash
# developer sweep (round 2f): vector-model kernels with the round-1 barrier sequence (default) vs the 9-10 barrier one
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02f
run() { echo "lib=${BISIP_B200_LIB:-default} $(timeout 120 python tools/kernel_time.py "$@" --reps 4 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["model"], "modes", j["n_modes"], "W", j["walkers"], "B", j["spectra"], "T", j["steps"], "%.3e" % j["evals_per_s"])')"; }
for rep in 1 2; do
for lib in "" bisip_b200/csrc/libbisip_b200_lr0.so; do
  if [ -n "$lib" ]; then export BISIP_B200_LIB=$lib; else unset BISIP_B200_LIB; fi
  run --model dias --walkers 128 --spectra 2368 --steps 500
  run --model dias --walkers 128 --spectra 1024 --steps 2000
  run --model dias --walkers 256 --spectra 1184 --steps 500
  run --model shin --walkers 128 --spectra 1776 --steps 500
  run --model shin --walkers 128 --spectra 1024 --steps 2000
  run --model colecole --n-modes 2 --walkers 128 --spectra 1776 --steps 500
done
done 2>&1 | tee gpurun_out/r02f/sweep_lr.log
