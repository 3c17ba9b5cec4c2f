#!/bin/bash
# developer sweep: collapsed decomposition kernel, CTAs/SM x proposals per thread
# (builds with -DBISIP_COLLAPSED_MINB=m -DBISIP_COLLAPSED_RPT=r -> libbisip_b200_c<m>r<r>.so; "default" = the shipped library)
cd "$(dirname "$0")/.."
O=gpurun_out/${ROUND:-r01h}; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_collapsed.py -x -q 2>&1 | tail -2
for v in default ${VARIANTS:-}; do
  lib=$PWD/bisip_b200/csrc/libbisip_b200_$v.so
  [ $v = default ] && lib=$PWD/bisip_b200/csrc/libbisip_b200.so
  [ -f $lib ] || continue
  echo "== $v"
  BISIP_B200_LIB=$lib python tools/kernel_time.py --model decomp --precision fp64-collapsed --B 3552 --W 256 --T 300 --reps 3
done 2>&1 | tee $O/collapsed_sweep2.log
