# per-phase cycle counters of CTA 0 (developer build, `make dbg`)
export BISIP_B200_LIB=$PWD/bisip_b200/csrc/libbisip_b200_dbg.so
python tools/kernel_time.py --model decomp --B 296 --W 256 --T 200 --reps 0
python tools/kernel_time.py --model dias --B 592 --W 128 --T 200 --reps 0
python tools/kernel_time.py --model shin --B 592 --W 128 --T 200 --reps 0
python tools/kernel_time.py --model colecole --K 2 --B 592 --W 128 --T 200 --reps 0
