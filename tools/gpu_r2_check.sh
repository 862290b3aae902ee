set -x
mkdir -p gpurun_out/g1
nvidia-smi -L > gpurun_out/g1/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g1/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/g1/pytest.log
tail -5 gpurun_out/g1/pytest.log
timeout 300 python tools/probe.py --chunks 16 1 64 129 1024 4096 8192 16384 32768 131072 > gpurun_out/g1/probe16.log 2>&1
tail -30 gpurun_out/g1/probe16.log
timeout 600 python bench.py > gpurun_out/g1/bench.json 2> gpurun_out/g1/bench.err; echo "bench exit $?"
tail -c 3000 gpurun_out/g1/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g1/bench_ref.json 2>&1
