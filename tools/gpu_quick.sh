out=gpurun_out/prof3; mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/bench_under_ncu.log 2>&1
BLSGPU_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_block129.csv \
    python tools/probe.py --chunks 4 129 > $out/probe129_under_ncu.log 2>&1
wc -l $out/*.csv; tail -3 $out/bench_under_ncu.log | cut -c1-300
