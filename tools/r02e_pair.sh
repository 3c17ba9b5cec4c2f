#!/bin/bash
# round 2e: pair-propose in the block-synchronous sampler (tcgen05 / DMMA / clustered evaluators) — parity + timings
cd "$(dirname "$0")/.."
O=gpurun_out/r02e; mkdir -p $O
( BISIP_SAMPLER=classic timeout 600 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_edges.py tests/test_gpu_collapsed.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 ) | tee $O/pytest_classic.log
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 ) | tee $O/pytest_default.log
run() { echo "sampler=${BISIP_SAMPLER:-default} $(timeout 120 python tools/kernel_time.py "$@" --steps 500 --reps 4 | tail -1 | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(j["model"], j["precision"], "W", j["walkers"], "N", j["n_freq"], "S", j["n_tau"], "B", j["spectra"], "%.3e" % j["evals_per_s"])')"; }
{
run --model decomp --precision 3xtf32 --spectra 592
run --model decomp --precision tf32 --spectra 592
run --model decomp --precision fp64 --n-tau 256 --spectra 296
run --model decomp --precision fp64 --n-tau 128 --spectra 296
run --model decomp --precision 3xtf32 --n-tau 256 --spectra 296
run --model decomp --precision tf32 --n-tau 256 --spectra 296
run --model decomp --precision fp64 --n-tau 32 --n-freq 32 --spectra 592
BISIP_SAMPLER=classic run --model decomp --precision fp64 --spectra 296
BISIP_SAMPLER=classic run --model decomp --precision fp64-collapsed --spectra 592
} 2>&1 | tee $O/timings.log
export BISIP_TIME_NOWARM=1 BISIP_SAMPLER=classic
san() { tool=$1; shift; echo "=== $tool $*"; compute-sanitizer --tool $tool --kernel-regex kns=ensemble_kernel python tools/kernel_time.py "$@" --reps 1 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|evals_per_s" | cut -c1-200 | head -6; }
{
san racecheck --model decomp --precision fp64 --spectra 2 --walkers 64 --steps 4
san racecheck --model decomp --precision fp64-collapsed --spectra 2 --walkers 33 --steps 4 --n-freq 20 --n-tau 40
san memcheck --model decomp --precision 3xtf32 --spectra 3 --walkers 64 --steps 6
} 2>&1 | tee $O/sanitizer.log
