# Round 1f: tolerance / throughput study of the reduced-precision decomposition kernels and one ncu --set full
# capture of the tcgen05 ensemble kernel (one GPU; run through gpurun, summaries copied to profiles/ afterwards).
set -x
O=gpurun_out/r01f
mkdir -p $O
python tools/tf32_study.py --out $O/r01f_tf32_study --spectra 296 --steps 1000 > $O/tf32_study.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -c 1 -f"
$NCU -o $O/ensemble_decomp_tcgen05 python tools/kernel_time.py --model decomp --B 296 --W 256 --T 200 --precision 3xtf32 --reps 0 > $O/ncu_tcgen05.log 2>&1
python tools/ncu_summary.py $O/ensemble_decomp_tcgen05.ncu-rep --name ensemble_decomp_tcgen05 --spectra 296 --scale-spectra 12500 \
  --note "round 1f, tcgen05 3xTF32 kernel, 296 spectra, W=256 T=200 N=64 S=64, two CTAs per SM" --out $O/r01f_ensemble_decomp_tcgen05 > /dev/null
for p in fp64 3xtf32-mma 3xtf32 tf32; do python tools/kernel_time.py --model decomp --B 592 --W 256 --T 500 --precision $p; done > $O/r01f_kernel_times.jsonl 2>&1
du -sh $O
