#!/bin/bash
# exchange kernel with the loads of 3 entries issued together vs one entry at a time (an experimental -DTB_XUNROLL build, measured here and then removed from the tree): parity subset, per-pass split on C3

cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -k "golden or mid_size or compact_stream or fused_pass or batch or api_surface" > $O/c33_tests.log 2>&1
echo "tests rc=$?" > $O/c33_box.log
for rep in 1 2; do
  timeout 300 python tools/probe_streams.py --workload c3 --passes 120 --rounds 1 > $O/c33_xu3_$rep.json 2> $O/c33_xu3_$rep.err
  TADBIT_B200_LIB=$PWD/tadbit_b200/variants/xu1.so timeout 300 python tools/probe_streams.py --workload c3 --passes 120 --rounds 1 > $O/c33_xu1_$rep.json 2> $O/c33_xu1_$rep.err
done
tail -3 $O/c33_tests.log; cat $O/c33_box.log
python - <<PY
import json
for f in ('c33_xu3_1','c33_xu1_1','c33_xu3_2','c33_xu1_2'):
    d=json.load(open('$O/%s.json' % f))
    for r in d['runs']: print(f, r['k1_streams'], 'ms/pass %.4f K1 %.4f exchange %.4f' % (r['ms_per_pass'], r['k1_ms'], r['exchange_ms']))
PY
