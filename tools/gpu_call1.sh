#!/bin/bash
# round-2 GPU call 1 (1 GPU): functional tests first, full-size parity, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi -L; free -g | head -2; nproc; } > $O/c1_box.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -k "not full_size and not batch_of_64" > $O/c1_tests_fast.log 2>&1
echo "fast tests rc=$?" >> $O/c1_box.log
timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -q -s -k "full_size or batch_of_64" > $O/c1_tests_full.log 2>&1
echo "full tests rc=$?" >> $O/c1_box.log
TB_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 > $O/c1_bench.json 2> $O/c1_bench.err
echo "bench rc=$?" >> $O/c1_box.log
tail -3 $O/c1_tests_fast.log; tail -5 $O/c1_tests_full.log; cat $O/c1_box.log; head -c 3000 $O/c1_bench.json; tail -20 $O/c1_bench.err
