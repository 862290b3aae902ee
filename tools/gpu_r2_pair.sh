set -x
out=gpurun_out/g2; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_golden_and_shares.py tests/test_gpu_batch_verify.py tests/test_gpu_io_and_verify.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -5 $out/pytest.log
python tools/probe.py --chunks 16 1024 2048 4096 8192 16384 32768 65536 131072 > $out/probe_pair_default.log 2>&1
BLSGPU_PAIR_HASH_MIN=100000000 BLSGPU_PAIR_LINES_MAX=0 python tools/probe.py --chunks 16 1024 2048 4096 8192 16384 32768 65536 131072 > $out/probe_pair_off.log 2>&1
BLSGPU_PAIR_HASH_MAX=100000000 BLSGPU_PAIR_LINES_MAX=100000000 python tools/probe.py --chunks 16 65536 131072 > $out/probe_pair_all.log 2>&1
BLSGPU_PAIR_HASH_MIN=1 BLSGPU_SMALL_LINES_MAX=0 python tools/probe.py --chunks 4 129 512 1024 2048 > $out/probe_pair_small.log 2>&1
grep -h "^n=\|hash" $out/probe_pair_default.log $out/probe_pair_off.log $out/probe_pair_all.log $out/probe_pair_small.log | cut -c1-250
python tools/probe_h2d.py > $out/h2d.log 2>&1; cat $out/h2d.log
ncu --set full --clock-control none --import-source on -k regex:k_rlc_scalars -s 2 -c 1 -f -o /tmp/prof_rlc python tools/probe.py --chunks 16 131072 > $out/ncu_rlc.log 2>&1
ncu -i /tmp/prof_rlc.ncu-rep --page details > $out/details_k_rlc_scalars.txt 2>&1
ncu -i /tmp/prof_rlc.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_src_summary.py /dev/stdin > $out/src_k_rlc_scalars.txt 2>&1
head -40 $out/src_k_rlc_scalars.txt
