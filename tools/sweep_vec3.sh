#!/bin/bash
# developer sweep (round 2c): frequencies in flight per thread (2 vs 4) after the branch-free loop
cd "$(dirname "$0")/.."
run() {
  local tag=$1; shift
  if [ -n "$tag" ]; then export BISIP_B200_LIB=bisip_b200/csrc/libbisip_b200_$tag.so; else unset BISIP_B200_LIB; fi
  echo "lib=${tag:-default} $(timeout 120 python tools/kernel_time.py "$@" --steps 500 --reps 3 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["model"], "modes", j["n_modes"], "W", j["walkers"], "N", j["n_freq"], "B", j["spectra"], "%.3e" % j["evals_per_s"])')"
}
for tag in "" ilp4 ilp4b; do
  run "$tag" --model dias --walkers 128 --spectra 2368
  run "$tag" --model dias --walkers 32 --n-freq 20 --spectra 9472
  run "$tag" --model dias --walkers 64 --spectra 4736
  run "$tag" --model colecole --n-modes 1 --walkers 128 --spectra 2368
  run "$tag" --model shin --walkers 128 --spectra 1776
  run "$tag" --model colecole --n-modes 2 --walkers 128 --spectra 1776
done
