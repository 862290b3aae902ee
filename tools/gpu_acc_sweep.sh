#!/bin/bash
for so in nim_blscurve_b200/libblsgpu.so build/variants/libblsgpu_acct2.so build/variants/libblsgpu_acct4.so; do
for gs in "16 2" "16 4" "32 1" "32 2" "32 4" "8 4" "64 1" "64 2"; do
  set -- $gs
  echo "=== $so G=$1 nseg=$2"
  BLSGPU_MILLER_G=$1 BLSGPU_MILLER_NSEG=$2 BLSGPU_LIB=$PWD/$so timeout 300 python tools/probe.py 131072 2>&1 | grep -o "rc=[-0-9]* wall=[0-9.]*ms\|miller_acc=[0-9.]*\|gt_product=[0-9.]*" | tr '\n' ' '; echo
done; done
