#!/bin/bash
# N4 (compartments): GPU parity tests, C2 timing of the DMMA X X^T and the eigen-solver, one full ncu capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_compartments.py -q --maxfail=20 > $O/c20_tests.log 2>&1
echo "tests rc=$?" > $O/c20_box.log
timeout 600 python tools/bench_compartments.py --workload c2 > $O/c20_cmp_c2.json 2> $O/c20_cmp_c2.err
echo "bench c2 rc=$?" >> $O/c20_box.log
timeout 600 python tools/bench_compartments.py --workload small --simple > $O/c20_cmp_small.json 2> $O/c20_cmp_small.err
echo "bench small rc=$?" >> $O/c20_box.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_syrk_dmma -c 1 -o $O/prof_syrk_c2 -f \
  python tools/bench_compartments.py --workload c2 --reps 1 --check 64 > $O/c20_ncu.log 2>&1
echo "ncu rc=$?" >> $O/c20_box.log
tail -25 $O/c20_tests.log; cat $O/c20_box.log; cat $O/c20_cmp_c2.json | head -c 3000; echo; tail -3 $O/c20_cmp_c2.err; cat $O/c20_cmp_small.json | head -c 1500
