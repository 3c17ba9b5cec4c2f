#!/bin/bash
# round 2d: block-synchronous sampler with 9 instead of 12 CTA barriers per step — parity, sanitizer, timings
cd "$(dirname "$0")/.."
O=gpurun_out/r02d; mkdir -p $O
( BISIP_SAMPLER=classic timeout 600 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_edges.py tests/test_gpu_collapsed.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 ) | tee $O/pytest_classic.log
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 ) | tee $O/pytest_default.log
run() { echo "sampler=${BISIP_SAMPLER:-default} $(timeout 120 python tools/kernel_time.py "$@" --steps 500 --reps 3 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["model"], j["precision"], "modes", j["n_modes"], "W", j["walkers"], "N", j["n_freq"], "S", j["n_tau"], "B", j["spectra"], "%.3e" % j["evals_per_s"])')"; }
{
run --model dias --walkers 128 --spectra 2368
run --model dias --walkers 256 --spectra 1184
run --model shin --walkers 128 --spectra 1776
run --model colecole --n-modes 2 --walkers 128 --spectra 1776
run --model decomp --precision 3xtf32 --spectra 592
run --model decomp --precision tf32 --spectra 592
run --model decomp --precision fp64 --n-tau 256 --spectra 296
run --model decomp --precision fp64 --n-tau 128 --spectra 296
run --model decomp --precision 3xtf32 --n-tau 256 --spectra 296
} 2>&1 | tee $O/timings.log
export BISIP_TIME_NOWARM=1 BISIP_SAMPLER=classic
san() { tool=$1; shift; echo "=== $tool $*"; compute-sanitizer --tool $tool --kernel-regex kns=ensemble_kernel python tools/kernel_time.py "$@" --reps 1 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|evals_per_s" | head -6; }
{
san racecheck --model dias --spectra 3 --walkers 128 --steps 4
san racecheck --model colecole --n-modes 2 --spectra 2 --walkers 33 --n-freq 20 --steps 5
san memcheck --model decomp --precision 3xtf32 --spectra 3 --walkers 64 --steps 6
san memcheck --model dias --spectra 3 --walkers 128 --steps 4
} 2>&1 | tee $O/sanitizer.log
