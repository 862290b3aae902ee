#!/bin/bash
# compute-sanitizer over one batch per route (warp-per-set programs, lane pairs, thread per set + deferred signature pair,
# sliced host copy), the MSM and the single-call entry points.  usage: tools/gpu_sanitize.sh <outdir under gpurun_out>
out=gpurun_out/${1:-san}; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # run <tool> <name> <cmd...>
  tool=$1; name=$2; shift 2
  timeout 1200 $CS --tool $tool --error-exitcode 9 --print-limit 20 "$@" > $out/${tool}_$name.log 2>&1
  echo "$tool $name exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tool}_$name.log | tail -1)"
}
export BLSGPU_GRAPH=0
run memcheck sizes python tools/sanitize_run.py --chunks 4 129 3000 9000 40001
run memcheck host_sliced python tools/sanitize_run.py --host --chunks 16 40001
run racecheck small python tools/sanitize_run.py --chunks 4 129 3000
run racecheck mid python tools/sanitize_run.py --chunks 4 9000
run initcheck sizes python tools/sanitize_run.py --chunks 4 129 3000 9000 40001
run synccheck sizes python tools/sanitize_run.py --chunks 4 129 3000 9000 40001
for f in $out/*.log; do echo "--- $f"; grep -h "^n=\|msm_g1\|aggregate\|SUMMARY" $f; done > $out/summary.txt
cat $out/summary.txt
