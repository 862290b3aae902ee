for i in 1 2; do
echo "== wait on"; python tools/probe.py --chunks 16 16384 32768 65536 131072 2>&1 | grep "^n=" | cut -c1-210
echo "== wait off"; BLSGPU_CHAIN_WAIT=0 python tools/probe.py --chunks 16 16384 32768 65536 131072 2>&1 | grep "^n=" | cut -c1-210
done
