#!/bin/bash
# run on a multi-GPU box (gpurun --gpus N): NCCL parity of the sharded batch (BASELINE configs[3], configs[4]) and the
# bench line of both arms at each world size.   usage: tools/gpu_multi.sh <outdir under gpurun_out> <world sizes...>
out=gpurun_out/$1; shift
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29600
for w in "$@"; do
  port=$((port + 1))
  timeout 600 $TR --nproc-per-node $w --master-port $port tools/nccl_parity.py 32768 > $out/nccl_parity_${w}gpu_32768.log 2>&1
  tail -1 $out/nccl_parity_${w}gpu_32768.log
  port=$((port + 1))
  timeout 600 $TR --nproc-per-node $w --master-port $port bench.py --gpus $w > $out/bench_${w}gpu.json 2> $out/bench_${w}gpu.err
  tail -c 400 $out/bench_${w}gpu.json; echo
done
w=${@: -1}
port=$((port + 1))
timeout 900 $TR --nproc-per-node $w --master-port $port tools/nccl_parity.py 1048576 > $out/nccl_parity_${w}gpu_1M.log 2>&1
tail -1 $out/nccl_parity_${w}gpu_1M.log
port=$((port + 1))
timeout 600 $TR --nproc-per-node $w --master-port $port bench.py --impl reference --gpus $w --steps 3 --warmup 1 > $out/bench_reference_${w}gpu.json 2>> $out/bench_${w}gpu.err
tail -c 300 $out/bench_reference_${w}gpu.json; echo
nvidia-smi topo -m > $out/topo.txt 2>&1
