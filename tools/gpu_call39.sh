#!/bin/bash
# N3 extraction side with the lines formatted by the library: GPU tests again, then the timing on C3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 100 python -m pytest tests/test_extraction.py -m gpu -q --maxfail=3 > $O/c39_tests.log 2>&1
echo "tests rc=$?" > $O/c39_box.log
timeout 120 python tools/bench_extraction.py --workload c3 --reps 2 > $O/c39_extract_c3.json 2> $O/c39_extract_c3.err
echo "bench rc=$?" >> $O/c39_box.log
tail -3 $O/c39_tests.log; cat $O/c39_box.log; tail -3 $O/c39_extract_c3.err; python - <<PY
import json
d=json.load(open('$O/c39_extract_c3.json'))
for r in d['runs']:
    print(r)
PY
