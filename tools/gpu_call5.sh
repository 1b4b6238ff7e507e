#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -k "not full_size and not batch_of_64" > $O/c5_tests_fast.log 2>&1
echo "fast tests rc=$?" > $O/c5_box.log
TB_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 > $O/c5_bench.json 2> $O/c5_bench.err
echo "bench rc=$?" >> $O/c5_box.log
# launch list of the same command (short run), after it exited 0 without ncu
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/c5_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > $O/c5_ncu_launch.log 2>&1
echo "launch list rc=$?" >> $O/c5_box.log
# DRAM traffic of one K1 evaluation (both kernels)
PROBE_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_rowsum -c 6 --csv --log-file $O/c5_traffic.csv python tools/probe_k1.py c3 2 > $O/c5_ncu_traffic.log 2>&1
echo "traffic rc=$?" >> $O/c5_box.log
python tools/run_onepass.py c3 > $O/c5_onepass.log 2>&1
tail -3 $O/c5_tests_fast.log; cat $O/c5_box.log; grep "bench e2e\|per pass\|sparse tiles" $O/c5_bench.err | tail -5; tail -8 $O/c5_traffic.csv | cut -c1-400; grep K5 $O/c5_onepass.log | tail -1
python - <<'PY'
import json
for line in open('gpurun_out/c5_bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','passes_per_step','converged','final_dev','ms_per_pass')})
        print(d['e2e']); print(d['cpu_baseline'])
PY
