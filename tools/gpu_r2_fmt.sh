set -x
out=gpurun_out/${1:-g7}; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -5 $out/pytest.log
BLSGPU_GRAPH=0 python tools/probe.py --chunks 4 1 129 1024 4096 16384 > $out/probe_f2.log 2>&1; grep -A1 "^n=" $out/probe_f2.log | cut -c1-260
BLSGPU_GRAPH=0 BLSGPU_PROG_FORMAT=1 python tools/probe.py --chunks 4 1 129 1024 4096 16384 > $out/probe_f1.log 2>&1; grep -A1 "^n=" $out/probe_f1.log | cut -c1-260
python tools/probe.py --chunks 4 1 64 129 512 1024 2047 > $out/probe_small.log 2>&1; grep "^n=" $out/probe_small.log
python tools/msm_probe.py 16 20 > $out/msm.log 2>&1; tail -5 $out/msm.log
