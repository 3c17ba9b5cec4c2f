for lib in libbisip_b200.so libbisip_b200_t128.so libbisip_b200_t128m6.so; do
echo "== $lib"
export BISIP_B200_LIB=$PWD/bisip_b200/csrc/$lib
python tools/kernel_time.py --model dias --B 3552 --W 128 --T 200
python tools/kernel_time.py --model shin --B 3552 --W 128 --T 200
python tools/kernel_time.py --model colecole --K 2 --B 3552 --W 128 --T 200
python tools/kernel_time.py --model colecole --K 1 --B 3552 --W 128 --T 200
done
