#!/bin/bash
# run on the GPU box: launch list of the bench command + one `ncu --set full` capture per hot kernel, exported to
# text/csv on the box (the .ncu-rep files are too large to bring back together).
# usage: tools/gpu_profile.sh <outdir under gpurun_out> [kernel regexes...]      env: PROBE="tools/probe.py 131072"
out=gpurun_out/$1; shift
probe=${PROBE:-tools/probe.py 131072}
mkdir -p $out
if [ -z "$NO_LAUNCH_LIST" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
fi
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-1} -c 1 -f -o /tmp/prof_$k \
      python $probe > $out/ncu_$k.log 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page details > $out/details_$k.txt 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > $out/raw_$k.csv 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_src_summary.py /dev/stdin > $out/src_$k.txt 2>&1
  rm -f /tmp/prof_$k.ncu-rep
done
ls -la $out
