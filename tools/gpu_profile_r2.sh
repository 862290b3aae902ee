#!/bin/bash
# Round-2 profile run on the GPU box: launch lists (bench, block batch) + one `ncu --set full` capture per kernel of
# interest, exported to text on the box (the .ncu-rep files are too large to bring back together).
# usage: tools/gpu_profile_r2.sh <outdir under gpurun_out>
out=gpurun_out/${1:-prof}; mkdir -p $out
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/bench_under_ncu.log 2>&1
BLSGPU_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_block129.csv \
    python tools/probe.py --chunks 4 129 > $out/probe129_under_ncu.log 2>&1
cap() {   # cap <kernel regex> <skip> <probe args...>
  k=$1; skip=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o /tmp/prof_$k python tools/probe.py "$@" > $out/ncu_$k.log 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page details > $out/details_$k.txt 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > $out/raw_$k.csv 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_src_summary.py /dev/stdin > $out/src_$k.txt 2>&1
  python tools/ncu_raw_pick.py $out/raw_$k.csv > $out/raw_pick_$k.txt 2>&1
  rm -f /tmp/prof_$k.ncu-rep
}
cap 'k_hash_sets$' 1 --chunks 16 131072
cap k_miller_lines$ 1 --chunks 16 131072
cap k_miller_acc_team 2 --chunks 16 131072
cap k_g1_mul 1 --chunks 16 131072
cap k_pairs_affine 1 --chunks 16 131072
cap k_hash_sets_lanes2 1 --chunks 16 16384
cap k_miller_lines_lanes2 1 --chunks 16 16384
cap k_hash_map_lanes2 1 --chunks 4 129
ls -la $out
