# Full one-GPU validation: GPU test suite, host-buffer probe, stage probe, bench line.  usage: bash tools/gpu_r2_full.sh <outdir>
set -x
out=gpurun_out/${1:-g4}; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -5 $out/pytest.log
python tools/probe_h2d.py > $out/h2d.log 2>&1; cat $out/h2d.log
python tools/probe.py --chunks 16 129 4096 16384 32768 131072 > $out/probe.log 2>&1; grep -A1 "^n=" $out/probe.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench exit $?"; tail -c 1500 $out/bench.err
python - <<PY
import json
d=json.loads(open('$out/bench.json').read().strip().splitlines()[-1])
print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','stages_ms','config3')},indent=1)[:3000])
print({k:v['batches_per_s'] for k,v in d['streaming_blocks']['runs'].items()}, d['streaming_blocks'].get('vs_cpu'))
print(d['block_batch']['ms'], d['config1']['chunks_4'], d['msm_g1']['value'])
print({k:round(v['ms'],2) for k,v in d['batch_sizes'].items()})
print({k:round(v['ms'],2) for k,v in d['chunk_sweep']['by_rlc_chunks'].items()})
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
