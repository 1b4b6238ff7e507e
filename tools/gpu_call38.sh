#!/bin/bash
# N3 extraction side timed on the genome-wide 5 kb workload (one GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 150 python tools/bench_extraction.py --workload c3 --reps 2 > $O/c38_extract_c3.json 2> $O/c38_extract_c3.err
echo "rc=$?" > $O/c38_box.log
cat $O/c38_box.log; tail -3 $O/c38_extract_c3.err; head -c 3000 $O/c38_extract_c3.json
