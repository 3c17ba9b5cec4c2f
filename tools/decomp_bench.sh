# parity suite + decomposition ensemble-kernel timings (+ per-phase cycles of the debug build)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/kernel_time.py --model decomp --B 592 --W 256 --T 300
python tools/kernel_time.py --model decomp --S 128 --B 296 --W 256 --T 100
python tools/kernel_time.py --model decomp --S 256 --B 296 --W 256 --T 100
python tools/kernel_time.py --model dias --B 3552 --W 128 --T 200
python tools/kernel_time.py --model shin --B 3552 --W 128 --T 200
python tools/kernel_time.py --model colecole --K 2 --B 3552 --W 128 --T 200
BISIP_B200_LIB=$PWD/bisip_b200/csrc/libbisip_b200_dbg.so python tools/kernel_time.py --model decomp --B 296 --W 256 --T 200 --reps 0 | grep -E "phase|fine"
