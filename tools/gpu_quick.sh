for m in 16384 2048; do
echo "== defer_min=$m"
BLSGPU_DEFER_SIG_MIN=$m python tools/probe.py --chunks 16 2048 4096 8192 16384 32768 2>&1 | grep "^n=" | cut -c1-215
done
BLSGPU_DEFER_SIG_MIN=2048 python -m pytest tests/test_gpu_golden_and_shares.py -m gpu -x -q 2>&1 | tail -2
