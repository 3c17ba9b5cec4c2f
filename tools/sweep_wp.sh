#!/bin/bash
# developer sweep: warp-private vs block-synchronous ("classic") sampler per model and walker count
for model in "dias" "shin" "colecole --n-modes 1" "colecole --n-modes 2" "decomp --precision fp64-collapsed"; do
  for wn in "32 20" "64 20" "64 64" "128 64" "256 64"; do
    set -- $wn; W=$1; N=$2
    S=$(( 148 * 2048 / W ))
    a=$(python tools/kernel_time.py --model $model --walkers $W --n-freq $N --n-tau $N --spectra $S --steps 300 --reps 4 | tail -1 | python -c 'import sys,json; print("%.3e" % json.loads(sys.stdin.read())["evals_per_s"])')
    b=$(BISIP_SAMPLER=classic python tools/kernel_time.py --model $model --walkers $W --n-freq $N --n-tau $N --spectra $S --steps 300 --reps 4 | tail -1 | python -c 'import sys,json; print("%.3e" % json.loads(sys.stdin.read())["evals_per_s"])')
    echo "$model W=$W N=$N spectra=$S  wp $a  classic $b"
  done
done
for cfg in "--model decomp --precision 3xtf32 --spectra 592" "--model decomp --precision tf32 --spectra 592" "--model decomp --precision fp64 --spectra 592"; do
  echo "classic $(python tools/kernel_time.py $cfg --steps 500 --reps 3 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["precision"], j["n_tau"], "%.3e" % j["evals_per_s"])')"
done
