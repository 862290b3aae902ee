set -x
out=gpurun_out/${1:-g8}; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_batch_verify.py tests/test_gpu_golden_and_shares.py tests/test_gpu_msm.py tests/test_gpu_io_and_verify.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -3 $out/pytest.log
BLSGPU_GRAPH=0 python tools/probe.py --chunks 4 1 129 1024 4096 > $out/probe_f2.log 2>&1; grep -A1 "^n=" $out/probe_f2.log | cut -c1-260
python tools/probe.py --chunks 4 1 64 129 512 1024 2047 > $out/probe_small.log 2>&1; grep "^n=" $out/probe_small.log
python tools/msm_probe.py 16 20 > $out/msm.log 2>&1; tail -3 $out/msm.log
