#!/bin/bash
# run on a multi-GPU box: BASELINE configs[3] — ONE 32 768-set batch per step sharded over 1/2/4/8 ranks (strong scaling)
out=gpurun_out/$1; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29700
python bench.py --gpus 1 --sets-per-gpu 32768 --no-msm --no-cpu-baseline --steps 20 --warmup 3 > $out/config3_1gpu.json 2> $out/config3.err
for w in 2 4 8; do
  port=$((port + 1))
  timeout 600 $TR --nproc-per-node $w --master-port $port bench.py --gpus $w --sets-per-gpu $((32768 / w)) --no-msm --no-cpu-baseline \
      --steps 20 --warmup 3 > $out/config3_${w}gpu.json 2>> $out/config3.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/*/config3_*gpu.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f.split("/")[-1], d["n_gpus"], "GPUs", "%.0f sets/s" % d["value"], "%.2f ms/step" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"])
PY
