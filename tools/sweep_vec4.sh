#!/bin/bash
# developer sweep (round 2c): warp-private vs block-synchronous sampler for the vector models on the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02c
run() {
  echo "sampler=${BISIP_SAMPLER:-default} $(timeout 120 python tools/kernel_time.py "$@" --steps 500 --reps 3 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["model"], "modes", j["n_modes"], "W", j["walkers"], "N", j["n_freq"], "B", j["spectra"], "%.3e" % j["evals_per_s"])')"
}
for smp in wp classic; do
  export BISIP_SAMPLER=$smp
  run --model dias --walkers 128 --spectra 2368
  run --model dias --walkers 256 --spectra 1184
  run --model shin --walkers 128 --spectra 1776
  run --model shin --walkers 256 --spectra 888
  run --model colecole --n-modes 1 --walkers 128 --spectra 2368
  run --model colecole --n-modes 2 --walkers 128 --spectra 1776
done
unset BISIP_SAMPLER
run --model dias --walkers 128 --spectra 1024
run --model shin --walkers 128 --spectra 1024
run --model colecole --n-modes 2 --walkers 64 --n-freq 20 --spectra 1024
