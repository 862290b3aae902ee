# Batch-size sweep with stage times (default thresholds, and [r_i]pk_i kept on the main stream) after the GPU suite.
set -x
out=gpurun_out/${1:-g9}; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -3 $out/pytest.log
python tools/probe.py --chunks 16 129 1024 1025 2048 4096 8192 16384 32768 65536 131072 > $out/probe.log 2>&1; grep -A1 "^n=" $out/probe.log | cut -c1-260
BLSGPU_G1_ASIDE_MAX=2047 python tools/probe.py --chunks 16 4096 16384 32768 > $out/probe_g1main.log 2>&1; grep -A1 "^n=" $out/probe_g1main.log | cut -c1-260
