#!/bin/bash
# run on the GPU box: stage timings for each prebuilt kernel variant.  usage: gpu_variants.sh "<sizes>" lib.so...
sizes=$1; shift
for so in "$@"; do
  echo "=== $so  G=$BLSGPU_MILLER_G nseg=$BLSGPU_MILLER_NSEG"
  BLSGPU_LIB=$PWD/$so timeout 300 python tools/probe.py $sizes 2>&1 | grep -v "^make_sets\|^imad"
done
