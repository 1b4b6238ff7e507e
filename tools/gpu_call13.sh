#!/bin/bash
# final single-GPU session of round 2: whole GPU test suite, smoke, bench, traffic + launch list of the final layout
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > $O/c13_tests_all.log 2>&1
echo "all gpu tests rc=$?" > $O/c13_box.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c13_smoke.log 2>&1
echo "smoke rc=$?" >> $O/c13_box.log
TB_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 > $O/c13_bench.json 2> $O/c13_bench.err
echo "bench rc=$?" >> $O/c13_box.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/c13_bench_ref.json 2> $O/c13_bench_ref.err
echo "bench reference rc=$?" >> $O/c13_box.log
PROBE_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_rowsum -c 4 --csv --log-file $O/c13_traffic.csv python tools/probe_k1.py c3 2 > $O/c13_ncu_traffic.log 2>&1
echo "traffic rc=$?" >> $O/c13_box.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/c13_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > $O/c13_ncu_launch.log 2>&1
echo "launch list rc=$?" >> $O/c13_box.log
PROBE_ONLY=1 timeout 900 ncu --set full --import-source on -k regex:k_rowsum -c 2 -o $O/prof_k1_c3_r02b python tools/probe_k1.py c3 2 > $O/c13_ncu_k1.log 2>&1
echo "ncu k1 rc=$?" >> $O/c13_box.log
timeout 300 python tools/bench_batch.py 64 > $O/c13_batch.json 2>&1
tail -4 $O/c13_tests_all.log; cat $O/c13_box.log; cat $O/c13_smoke.log | tail -2; grep "bench e2e\|per pass\|CTA 0" $O/c13_bench.err | tail -4; cat $O/c13_batch.json | tail -1 | cut -c1-600
python - <<'PY'
import json
for f in ('gpurun_out/c13_bench.json','gpurun_out/c13_bench_ref.json'):
  for line in open(f):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d.get(k) for k in ('impl','value','ms_per_step','passes_per_step','converged','final_dev','ms_per_pass')})
        print(d.get('e2e')); print(d.get('cpu_baseline')); print(d.get('reference_python'))
        if 'roofline' in d: print({k:d['roofline'][k] for k in ('frac','frac_algorithmic','traffic','kernel_ms','kernels')}); print(d['pipeline'])
PY
