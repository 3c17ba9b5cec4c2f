#!/bin/bash
# developer sweep (round 2b): vector-model kernels, alternative builds (tools/build_variant.sh) on the same configs
cd "$(dirname "$0")/.."
run() {  # lib-tag, config...
  local tag=$1; shift
  if [ -n "$tag" ]; then export BISIP_B200_LIB=bisip_b200/csrc/libbisip_b200_$tag.so; else unset BISIP_B200_LIB; fi
  echo "lib=${tag:-default} $(timeout 120 python tools/kernel_time.py "$@" --steps 500 --reps 3 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["model"], "modes", j["n_modes"], "W", j["walkers"], "N", j["n_freq"], "B", j["spectra"], "%.3e" % j["evals_per_s"], "acc %.3f" % j["acceptance"])')"
}
DIAS="--model dias --walkers 128 --spectra 2368"
SHIN="--model shin --walkers 128 --spectra 1776"
CC1="--model colecole --n-modes 1 --walkers 128 --spectra 2368"
CC2="--model colecole --n-modes 2 --walkers 128 --spectra 1776"
CC2S="--model colecole --n-modes 2 --walkers 64 --n-freq 20 --spectra 3552"
DIASS="--model dias --walkers 32 --n-freq 20 --spectra 9472"
for tag in base "" u2; do
  for cfg in "$DIAS" "$SHIN" "$CC1" "$CC2" "$CC2S" "$DIASS"; do run "$tag" $cfg; done
done
for cfg in "$SHIN" "$CC1" "$CC2" "$CC2S"; do run noexptab $cfg; done
for cfg in "$SHIN" "$CC2" "$CC2S"; do run shin64 $cfg; done
