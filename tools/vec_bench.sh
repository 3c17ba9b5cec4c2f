# GPU parity tests + ensemble-kernel timings of every model at the shapes quoted in profiles/
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/kernel_time.py --model dias --B 3552 --W 128 --T 200
python tools/kernel_time.py --model shin --B 3552 --W 128 --T 200
python tools/kernel_time.py --model colecole --K 2 --B 3552 --W 128 --T 200
python tools/kernel_time.py --model colecole --K 1 --B 3552 --W 128 --T 200
python tools/kernel_time.py --model dias --B 1776 --W 256 --T 200
python tools/kernel_time.py --model decomp --B 592 --W 256 --T 300
python tools/kernel_time.py --model decomp --S 128 --B 296 --W 256 --T 100
python tools/kernel_time.py --model decomp --S 256 --B 296 --W 256 --T 100
