#!/bin/bash
# compute-sanitizer over the round-2 kernels (small shapes; under gpurun)
run() { tool=$1; shift; echo "=== $tool $*"; compute-sanitizer --tool $tool --kernel-regex kns=$KRE python tools/kernel_time.py "$@" --reps 1 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|evals_per_s" | head -6; }
export BISIP_TIME_NOWARM=1
KRE=ensemble_wp
for tool in memcheck racecheck; do
  run $tool --model decomp --precision fp64-collapsed --spectra 3 --walkers 64 --steps 6
  run $tool --model decomp --precision fp64-collapsed --spectra 2 --walkers 33 --steps 5 --poly-deg 9 --n-freq 20 --n-tau 40
  run $tool --model decomp --precision fp64 --spectra 2 --walkers 256 --steps 4
  run $tool --model colecole --n-modes 2 --spectra 3 --walkers 64 --n-freq 20 --steps 6
  run $tool --model dias --spectra 3 --walkers 31 --n-freq 17 --steps 6
done
KRE=column_stats
compute-sanitizer --tool memcheck --kernel-regex kns=column_stats python tools/stats_bench.py --spectra 3 --n 25600 2>&1 | grep -E "ERROR SUMMARY|max" | head -3
compute-sanitizer --tool racecheck --kernel-regex kns=column_stats python tools/stats_bench.py --spectra 3 --n 25600 2>&1 | grep -E "RACECHECK SUMMARY|hazard|max" | head -5
KRE=ensemble_kernel
export BISIP_SAMPLER=classic
run memcheck --model decomp --precision 3xtf32 --spectra 3 --walkers 64 --steps 6
run racecheck --model dias --spectra 3 --walkers 128 --steps 4
