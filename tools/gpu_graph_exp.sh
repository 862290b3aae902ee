BLSGPU_GRAPH_MAX=8192 python tools/probe.py --chunks 16 2048 4096 8192 2>&1 | grep "^n="
python tools/probe.py --chunks 16 2048 4096 8192 2>&1 | grep "^n="
BLSGPU_GRAPH_MAX=8192 python -m pytest tests/test_gpu_golden_and_shares.py -m gpu -x -q -k "route_boundaries or large_batch" 2>&1 | tail -2
python -c "
import __graft_entry__ as g
g.smoke()
"
