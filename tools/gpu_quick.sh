for i in 1 2; do
echo "== with fp_sqr_n"; python tools/probe.py --chunks 16 131072 2>&1 | grep -A1 "^n=" | cut -c1-200
echo "== without"; BLSGPU_LIB=$PWD/build/libblsgpu_nosqrn.so python tools/probe.py --chunks 16 131072 2>&1 | grep -A1 "^n=" | cut -c1-200
done
echo "== hog off"; BLSGPU_CHAIN_HOG_MIN=100000000 python tools/probe.py --chunks 16 131072 2>&1 | grep -A1 "^n=" | cut -c1-200
echo "== chunks 1024"; python tools/probe.py --chunks 1024 131072 2>&1 | grep -A1 "^n=" | cut -c1-200
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu --format=csv
