#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
V=tadbit_b200/variants/rows4096.so
TADBIT_B200_LIB=$V timeout 600 python tools/kernel_coverage.py > $O/c12_cov_rows4096.log 2>&1
echo "coverage rows4096 rc=$?" > $O/c12_box.log
for lib in "" $V; do
  echo "== lib=${lib:-default}" >> $O/c12_probe.log
  TADBIT_B200_LIB=$lib PROBE_ONLY=1 timeout 300 python tools/probe_k1.py c3 20 2>&1 | grep "K1 ms\|tile_units" | cut -c1-300 >> $O/c12_probe.log
  for sh in 0/8 3/8 1/4; do
    TADBIT_B200_LIB=$lib PROBE_SHARD=$sh timeout 300 python tools/probe_k1.py c3 20 2>&1 | grep "K1 ms\|shard" >> $O/c12_probe.log
  done
done
timeout 600 python -m pytest tests -m gpu -q --maxfail=12 -k "windows or pairs or decay" > $O/c12_tests.log 2>&1
echo "new tests rc=$?" >> $O/c12_box.log
cat $O/c12_box.log; tail -2 $O/c12_cov_rows4096.log; cat $O/c12_probe.log; tail -3 $O/c12_tests.log
