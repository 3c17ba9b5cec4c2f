# ncu captures behind profiles/${ROUND:-r01f}_* (one GPU; run through gpurun).  Reports are summarised on the box
# (tools/ncu_summary.py) and only the decomposition report is kept: gpurun_out/ travels back only if <= 64 MiB.
set -x
O=gpurun_out/${ROUND:-r01f}
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -c 1 -f"
cap() {   # name spectra scale note -- kernel_time args
  name=$1; spectra=$2; scale=$3; note=$4; shift 4
  $NCU -o $O/$name python tools/kernel_time.py "$@" --reps 0 > $O/ncu_$name.log 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep --name $name --spectra $spectra --scale-spectra $scale --note "$note" --out $O/${ROUND:-r01f}_$name > /dev/null
}
cap ensemble_decomp 296 12500 "round 1f, 296 spectra, W=256 T=200 N=64 S=64" --model decomp --B 296 --W 256 --T 200
cap ensemble_decomp_rc256 296 10000 "round 1f, 296 spectra, W=256 T=60 N=64 S=256, 2-CTA cluster" --model decomp --S 256 --B 296 --W 256 --T 60
rm -f $O/ensemble_decomp_rc256.ncu-rep
cap ensemble_dias 1184 1024 "round 1f, 1184 spectra, W=128 T=200 N=64, 128-thread CTAs x 8/SM" --model dias --B 1184 --W 128 --T 200
rm -f $O/ensemble_dias.ncu-rep
cap ensemble_shin 888 1024 "round 1f, 888 spectra, W=128 T=200 N=64, 128-thread CTAs x 6/SM" --model shin --B 888 --W 128 --T 200
rm -f $O/ensemble_shin.ncu-rep
cap ensemble_colecole2 888 1024 "round 1f, 888 spectra, n_modes=2, W=128 T=200 N=64, 128-thread CTAs x 6/SM" --model colecole --K 2 --B 888 --W 128 --T 200
rm -f $O/ensemble_colecole2.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${ROUND:-r01f}_launches.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1
python bench.py > $O/${ROUND:-r01f}_bench.json 2> $O/bench.err
du -sh gpurun_out
