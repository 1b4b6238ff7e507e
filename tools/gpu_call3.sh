#!/bin/bash
# round-2 GPU call 3 (1 GPU): re-test, K5 profile, tile-fill variant, small-shard K1, ncu of K1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -k "not full_size and not batch_of_64" > $O/c3_tests_fast.log 2>&1
echo "fast tests rc=$?" > $O/c3_box.log
timeout 300 python tools/run_onepass.py c3s > $O/c3_onepass.log 2>&1
timeout 600 ncu --set full --import-source on -k regex:k_visit -c 2 -o $O/prof_k5_c3s python tools/run_onepass.py c3s > $O/c3_ncu_k5.log 2>&1
echo "ncu k5 rc=$?" >> $O/c3_box.log
# tile fill variant: build stages and K1 time, then the coverage script under the variant
TB_TIMING=1 PROBE_ONLY=1 timeout 300 python tools/probe_k1.py c3 > $O/c3_probe_default.log 2>&1
TB_TIMING=1 PROBE_ONLY=1 TADBIT_B200_LIB=tadbit_b200/variants/fillhist.so timeout 300 python tools/probe_k1.py c3 > $O/c3_probe_fillhist.log 2>&1
TADBIT_B200_LIB=tadbit_b200/variants/fillhist.so timeout 600 python tools/kernel_coverage.py > $O/c3_cov_fillhist.log 2>&1
echo "coverage fillhist rc=$?" >> $O/c3_box.log
# K1 of a 1/8 shard on one GPU, units per SM
for u in 16 10 6 4; do
  for sh in 0/8 3/8 7/8; do
    echo "== units $u shard $sh" >> $O/c3_shard.log
    TB_TILE_UNITS_PER_SM=$u PROBE_SHARD=$sh timeout 300 python tools/probe_k1.py c3 20 2>&1 | grep "K1 ms\|shard" >> $O/c3_shard.log
  done
done
# full ncu of the two K1 kernels on C3
PROBE_ONLY=1 timeout 900 ncu --set full --import-source on -k regex:k_rowsum -c 2 -o $O/prof_k1_c3_r02 python tools/probe_k1.py c3 2 > $O/c3_ncu_k1.log 2>&1
echo "ncu k1 rc=$?" >> $O/c3_box.log
tail -3 $O/c3_tests_fast.log; cat $O/c3_box.log; cat $O/c3_onepass.log; grep "sparse tiles\|K1 ms\|encode" $O/c3_probe_default.log $O/c3_probe_fillhist.log; tail -2 $O/c3_cov_fillhist.log; cat $O/c3_shard.log
