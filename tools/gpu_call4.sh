#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -k "not full_size and not batch_of_64" > $O/c4_tests_fast.log 2>&1
echo "fast tests rc=$?" > $O/c4_box.log
timeout 300 python tools/run_onepass.py c3 > $O/c4_onepass_c3.log 2>&1
TB_TIMING=1 PROBE_ONLY=1 TADBIT_B200_LIB=tadbit_b200/variants/fillhist.so timeout 300 python tools/probe_k1.py c3 > $O/c4_probe_fillhist.log 2>&1
TADBIT_B200_LIB=tadbit_b200/variants/fillhist.so timeout 600 python tools/kernel_coverage.py > $O/c4_cov_fillhist.log 2>&1
echo "coverage fillhist rc=$?" >> $O/c4_box.log
TB_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 > $O/c4_bench.json 2> $O/c4_bench.err
echo "bench rc=$?" >> $O/c4_box.log
tail -4 $O/c4_tests_fast.log; cat $O/c4_box.log; cat $O/c4_onepass_c3.log; grep "sparse tiles\|K1 ms\|encode" $O/c4_probe_fillhist.log; tail -2 $O/c4_cov_fillhist.log; grep "bench e2e\|per pass" $O/c4_bench.err | tail -4
python - <<'PY'
import json
for line in open('gpurun_out/c4_bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','passes_per_step','converged','final_dev','ms_per_pass')})
        print(d['e2e']); print(d['pipeline']); print(d['roofline']['frac'], d['roofline']['kernels'])
PY
