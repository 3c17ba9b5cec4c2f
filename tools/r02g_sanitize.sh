#!/bin/bash
# compute-sanitizer on the final build: the warp-private vector-model kernels after the round-2c loop / exp-table changes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02g
run() { tool=$1; shift; echo "=== $tool $*"; timeout 200 compute-sanitizer --tool $tool --kernel-regex kns=$KRE python tools/kernel_time.py "$@" --reps 1 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|evals_per_s" | cut -c1-160 | head -5; }
export BISIP_TIME_NOWARM=1
KRE=ensemble_wp
{
for tool in memcheck racecheck; do
  run $tool --model colecole --n-modes 2 --spectra 3 --walkers 64 --n-freq 20 --steps 6
  run $tool --model colecole --n-modes 1 --spectra 2 --walkers 128 --steps 4
  run $tool --model dias --spectra 3 --walkers 31 --n-freq 17 --steps 6
  run $tool --model shin --spectra 2 --walkers 64 --n-freq 20 --steps 5
done
} 2>&1 | tee gpurun_out/r02g/sanitizer_wp.log
