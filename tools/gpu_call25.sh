#!/bin/bash
# final single-GPU pass of round 2: all GPU tests, smoke, default bench, ncu launch list of a bench run
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > $O/c25_tests.log 2>&1
echo "tests rc=$?" > $O/c25_box.log
timeout 300 python __graft_entry__.py --smoke > $O/c25_smoke.log 2>&1
echo "smoke rc=$?" >> $O/c25_box.log
timeout 900 python bench.py > $O/c25_bench.json 2> $O/c25_bench.err
echo "bench rc=$?" >> $O/c25_box.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4400 --csv --log-file $O/c25_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > $O/c25_ncu_launch.log 2>&1
echo "ncu rc=$?" >> $O/c25_box.log
tail -5 $O/c25_tests.log; tail -2 $O/c25_smoke.log; cat $O/c25_box.log
python - <<PY
import json
for line in open('$O/c25_bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','passes_per_step','converged','final_dev','ms_per_pass','gpu_launches','clocks')})
        r=d['roofline']; print({k:r[k] for k in ('achieved','frac','frac_algorithmic','kernel_ms','traffic','kernels')})
        print(d['pipeline']); print(d['e2e']); print(d['cpu_baseline']); print(d['parity'])
PY
tail -3 $O/c25_bench.err
