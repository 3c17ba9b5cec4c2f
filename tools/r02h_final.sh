#!/bin/bash
# Round-2h check of the last commit on one B200 (under gpurun): GPU suite, smoke, the default bench line.
cd "$(dirname "$0")/.."
O=gpurun_out/r02h; mkdir -p $O
( time timeout 200 python -m pytest tests -m gpu -x -q ) > $O/r02h_pytest_gpu.log 2>&1; tail -4 $O/r02h_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02h_smoke.log 2>&1; tail -1 $O/r02h_smoke.log
timeout 400 python bench.py > $O/r02h_bench_n1.json 2> $O/bench.err; python tools/bench_digest.py $O/r02h_bench_n1.json
