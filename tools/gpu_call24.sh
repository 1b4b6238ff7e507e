#!/bin/bash
# Lanczos without per-step round trips: parity tests, timing on small chromosomes and on chr1 @10 kb
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_compartments.py -q --maxfail=20 > $O/c24_tests.log 2>&1
echo "tests rc=$?" > $O/c24_box.log
timeout 600 python tools/bench_compartments.py --workload small --reps 3 > $O/c24_cmp_small.json 2> $O/c24_cmp_small.err
echo "bench small rc=$?" >> $O/c24_box.log
timeout 600 python tools/bench_compartments.py --workload c2 --reps 2 > $O/c24_cmp_c2.json 2> $O/c24_cmp_c2.err
echo "bench c2 rc=$?" >> $O/c24_box.log
tail -4 $O/c24_tests.log; cat $O/c24_box.log
python - <<PY
import json
for f in ('$O/c24_cmp_small.json', '$O/c24_cmp_c2.json'):
    d=json.load(open(f))
    for r in d['runs']: print(r['good_bins'], r['gemm'], 'gemm %.3f ms %.2f TF/s' % (r['timing']['gemm_ms'], r['tflops']), 'eig %.3f ms wall %.3f' % (r['timing']['eig_ms'], r['eig_wall_ms']), r['lanczos'], r['eigenvalues'])
    print(d['check'])
PY
