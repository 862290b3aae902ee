#!/bin/bash
# run on the GPU box: Miller accumulation shape (pairs per group, loop segments) for small batches
for gs in "1 32" "2 32" "4 16" "4 32" "8 8" "8 16" "16 4" "16 8" "16 16" "32 4" "32 8" "64 4" "130 2" "130 8"; do
  set -- $gs
  echo "=== G=$1 nseg=$2"
  BLSGPU_MILLER_G=$1 BLSGPU_MILLER_NSEG=$2 timeout 300 python tools/probe.py ${SIZES:-129 1024} 2>&1 | grep -o "^n=[0-9]* rc=[-0-9]* wall=[0-9.]*ms\|miller_acc=[0-9.]*\|gt_product=[0-9.]*\|partial=[0-9.]*" | tr '\n' ' '; echo
done
