for L in 8 16 32; do echo "== L=$L"; BLSGPU_MSM_L=$L python tools/msm_probe.py 16 20 2>&1 | tail -2; done
for G in 2 4; do echo "== GROUPS=$G"; BLSGPU_MSM_GROUPS=$G python tools/msm_probe.py 16 20 2>&1 | tail -2; done
for K in 16 24 48; do echo "== K=$K"; BLSGPU_MSM_K=$K python tools/msm_probe.py 20 2>&1 | tail -1; done
for C in 14 15; do echo "== C=$C"; BLSGPU_MSM_C=$C python tools/msm_probe.py 20 2>&1 | tail -1; done
