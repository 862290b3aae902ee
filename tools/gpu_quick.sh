timeout 900 python -m pytest tests/test_gpu_golden_and_shares.py -m gpu -x -q -k "epoch" 2>&1 | tail -5
python tools/probe.py --chunks 16 131072 2>&1 | grep -A1 "^n=" | cut -c1-230
