#!/bin/bash
# K5 with 32-bit shared counters (parity + timing on C3), K1 of 1/8 shards serial vs two streams, coverage under the checked build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -k "golden or coverage or api_surface or store or pairs_decay or windows" > $O/c22_tests.log 2>&1
echo "tests rc=$?" > $O/c22_box.log
timeout 600 python tools/run_onepass.py c3 > $O/c22_onepass_c3.log 2>&1
echo "onepass rc=$?" >> $O/c22_box.log
timeout 900 python tools/probe_shard.py --workload c3 --world 8 --ranks 0,3,7 > $O/c22_shard8.json 2> $O/c22_shard8.err
echo "shard8 rc=$?" >> $O/c22_box.log
timeout 900 python tools/probe_shard.py --workload c3 --world 2 --ranks 0 > $O/c22_shard2.json 2> $O/c22_shard2.err
echo "shard2 rc=$?" >> $O/c22_box.log
tail -4 $O/c22_tests.log; cat $O/c22_box.log; cat $O/c22_onepass_c3.log; cat $O/c22_shard8.json; echo; cat $O/c22_shard2.json; tail -3 $O/c22_shard8.err
