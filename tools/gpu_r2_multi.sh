# Multi-GPU validation on one box: NCCL parity + the multi-device C consumer + bench at N ranks + the reference arm.
# usage (through gpurun --gpus N): bash tools/gpu_r2_multi.sh N
set -x
N=${1:-8}
out=gpurun_out/m$N; mkdir -p $out
nvidia-smi -L > $out/smi.txt; nproc >> $out/smi.txt
timeout 1200 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_abi_multi.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -4 $out/pytest.log
./tests/c_abi_multi $((131072 * N)) 0 > $out/c_abi_multi_full.log 2>&1; cat $out/c_abi_multi_full.log
./tests/c_abi_multi 32768 0 > $out/c_abi_multi_32768.log 2>&1; cat $out/c_abi_multi_32768.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench exit $?"; tail -c 600 $out/bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $out/bench_ref.json 2>&1
python - <<PY
import json
d=json.loads([l for l in open('$out/bench.json') if l.startswith('{')][-1])
print(json.dumps({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e','config3','parity_checks','stages_ms','config')},indent=1)[:3000])
PY
