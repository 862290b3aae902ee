set -x
N=2
out=gpurun_out/m2b; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -5 $out/pytest.log
python tools/probe.py --chunks 4 1 64 129 512 1024 2047 > $out/probe_small.log 2>&1; grep "^n=" $out/probe_small.log
BLSGPU_GRAPH=0 python tools/probe.py --chunks 4 129 > $out/probe_small_ng.log 2>&1; grep -A1 "^n=" $out/probe_small_ng.log
BLSGPU_GRAPH=0 BLSGPU_MAP_LANES2=0 python tools/probe.py --chunks 4 129 > $out/probe_small_ng_oldmap.log 2>&1; grep -A1 "^n=" $out/probe_small_ng_oldmap.log
python tools/probe_h2d.py > $out/h2d.log 2>&1; cat $out/h2d.log
./tests/c_abi_multi 131072 0 > $out/c_abi_multi_131072.log 2>&1; cat $out/c_abi_multi_131072.log
./tests/c_abi_multi 32768 0 > $out/c_abi_multi_32768.log 2>&1; cat $out/c_abi_multi_32768.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench exit $?"; tail -c 800 $out/bench.err
python - <<PY
import json
d=json.loads([l for l in open('$out/bench.json') if l.startswith('{')][-1])
print(json.dumps({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e','config3','parity_checks','stages_ms')},indent=1)[:2500])
PY
