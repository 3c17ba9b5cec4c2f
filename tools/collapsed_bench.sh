#!/bin/bash
# developer sweep: collapsed decomposition kernel at several CTAs/SM x rows per thread
# (builds with -DBISIP_COLLAPSED_MINB=m -DBISIP_COLLAPSED_RPT=r -> libbisip_b200_c<m>r<r>.so), then one ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out/${ROUND:-r01h}
mkdir -p $O
for v in c4r1 c4r2 c3r2 c3r4 c4r4; do
  lib=$PWD/bisip_b200/csrc/libbisip_b200_$v.so
  [ -f $lib ] || continue
  echo "== $v"
  BISIP_B200_LIB=$lib python tools/kernel_time.py --model decomp --precision fp64-collapsed --B 2368 --W 256 --T 500 --reps 2
done 2>&1 | tee $O/collapsed_sweep.log
NCU="ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -c 1 -f"
for v in ${NCU_VARIANTS:-c4r2}; do
  BISIP_B200_LIB=$PWD/bisip_b200/csrc/libbisip_b200_$v.so $NCU -o $O/collapsed_$v python tools/kernel_time.py --model decomp --precision fp64-collapsed --B 592 --W 256 --T 200 --reps 0 > $O/ncu_collapsed_$v.log 2>&1
  python tools/ncu_summary.py $O/collapsed_$v.ncu-rep --name ensemble_decomp_collapsed_$v --spectra 592 --scale-spectra 12500 --note "round 1h, 592 spectra, W=256 T=200 N=64 S=64, build $v" --out $O/${ROUND:-r01h}_ensemble_decomp_collapsed_$v > /dev/null
done
du -sh gpurun_out
