#!/bin/bash
# run on the GPU box: MSM parity tests + timing for window-group / chunk-size settings
for G in 1 2 4; do for K in 0 24 40; do
  echo "=== groups=$G K=$K"
  BLSGPU_MSM_GROUPS=$G BLSGPU_MSM_K=$K timeout 300 python tools/msm_probe.py ${SIZES:-20} 2>&1 | tail -n +1
done; done
