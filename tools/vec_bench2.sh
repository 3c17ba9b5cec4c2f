python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for lib in libbisip_b200.so libbisip_b200_m3.so; do
echo "== $lib"
export BISIP_B200_LIB=$PWD/bisip_b200/csrc/$lib
python tools/kernel_time.py --model dias --B 1776 --W 128 --T 200
python tools/kernel_time.py --model shin --B 1776 --W 128 --T 200
python tools/kernel_time.py --model colecole --K 2 --B 1776 --W 128 --T 200
python tools/kernel_time.py --model colecole --K 1 --B 1776 --W 128 --T 200
done
