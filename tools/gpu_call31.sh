#!/bin/bash
# tile kernel with a scheduler warp (an experimental -DTB_TILE_PREFETCH build that was measured here and then removed from the tree) vs the build without it: parity subset, K1 on C3 and on a 1/8 shard
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -k "golden or coverage or mid_size or compact_stream or fused_pass or store" > $O/c31_tests.log 2>&1
echo "tests rc=$?" > $O/c31_box.log
for rep in 1 2; do
  timeout 300 python tools/probe_k1.py c3 20 > $O/c31_pf_c3_$rep.log 2>&1
  TADBIT_B200_LIB=$PWD/tadbit_b200/variants/nopf.so timeout 300 python tools/probe_k1.py c3 20 > $O/c31_nopf_c3_$rep.log 2>&1
done
PROBE_SHARD=3/8 TB_TIMING=1 timeout 300 python tools/probe_k1.py c3 20 > $O/c31_pf_shard.log 2>&1
PROBE_SHARD=3/8 TB_TIMING=1 TADBIT_B200_LIB=$PWD/tadbit_b200/variants/nopf.so timeout 300 python tools/probe_k1.py c3 20 > $O/c31_nopf_shard.log 2>&1
tail -3 $O/c31_tests.log; cat $O/c31_box.log
for f in $O/c31_pf_c3_1.log $O/c31_nopf_c3_1.log $O/c31_pf_c3_2.log $O/c31_nopf_c3_2.log $O/c31_pf_shard.log $O/c31_nopf_shard.log; do echo "== $f"; grep -h "K1 ms\|normalize_hic\|k_rowsum_tiles:" $f; done
