set -x
out=gpurun_out/g3; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_batch_verify.py tests/test_gpu_golden_and_shares.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -5 $out/pytest.log
python tools/probe_h2d.py > $out/h2d.log 2>&1; cat $out/h2d.log
python tools/probe.py --chunks 4 1 64 129 512 1024 2047 > $out/probe_graph.log 2>&1
BLSGPU_GRAPH=0 python tools/probe.py --chunks 4 1 64 129 512 1024 2047 > $out/probe_nograph.log 2>&1
grep -h "^n=" $out/probe_graph.log $out/probe_nograph.log
timeout 600 python bench.py --no-msm > $out/bench.json 2> $out/bench.err; echo "bench exit $?"; tail -c 1500 $out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g3/bench.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','block_batch','streaming_blocks','config1','batch_sizes')},indent=1)[:6000])
PY
