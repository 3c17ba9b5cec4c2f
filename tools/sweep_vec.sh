#!/bin/bash
# developer sweep: time the vector-model / collapsed kernels with alternative builds of one translation unit
for lib in "" _d4r80 _d2r80 _d1r64; do
  export BISIP_B200_LIB=bisip_b200/csrc/libbisip_b200$lib.so
  for cfg in "--model dias --walkers 128 --spectra 2368" "--model shin --walkers 128 --spectra 1776"; do
    echo "lib=$lib $(python tools/kernel_time.py $cfg --steps 500 --reps 3 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["model"], j["walkers"], "%.3e" % j["evals_per_s"])')"
  done
done
for lib in "" _c80; do
  export BISIP_B200_LIB=bisip_b200/csrc/libbisip_b200$lib.so
  echo "lib=$lib $(python tools/kernel_time.py --model decomp --precision fp64-collapsed --steps 500 --reps 3 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["precision"], j["walkers"], "%.3e" % j["evals_per_s"])')"
done
unset BISIP_B200_LIB
for cfg in "--model decomp --precision 3xtf32 --spectra 592" "--model decomp --precision tf32 --spectra 592" "--model decomp --precision fp64 --spectra 592" "--model decomp --precision fp64 --n-tau 256 --spectra 296"; do
  echo "classic-int-draws $(python tools/kernel_time.py $cfg --steps 500 --reps 3 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["precision"], j["n_tau"], "%.3e" % j["evals_per_s"])')"
done
