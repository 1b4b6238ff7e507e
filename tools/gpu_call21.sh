#!/bin/bash
# banded tile order of the DMMA X X^T (DRAM traffic), K1 on two streams vs back to back on C3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_compartments.py -q --maxfail=20 > $O/c21_tests.log 2>&1
echo "tests rc=$?" > $O/c21_box.log
timeout 600 python tools/bench_compartments.py --workload c2 --reps 2 > $O/c21_cmp_c2.json 2> $O/c21_cmp_c2.err
echo "bench c2 rc=$?" >> $O/c21_box.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:k_syrk_dmma -c 1 --csv --log-file $O/c21_syrk_traffic.csv \
  python tools/bench_compartments.py --workload c2 --reps 1 --check 64 > $O/c21_ncu.log 2>&1
echo "ncu rc=$?" >> $O/c21_box.log
timeout 900 python tools/probe_streams.py --workload c3 --passes 60 --rounds 3 > $O/c21_streams_c3.json 2> $O/c21_streams_c3.err
echo "streams rc=$?" >> $O/c21_box.log
tail -4 $O/c21_tests.log; cat $O/c21_box.log
python - <<PY
import json
d=json.load(open('$O/c21_cmp_c2.json'))
for r in d['runs']: print(r['gemm'], r['timing']['gemm_ms'], r['tflops'], r['lanczos'])
print(d['check'])
d=json.load(open('$O/c21_streams_c3.json'))
for r in d['runs']: print(r)
PY
grep -v "^==" $O/c21_syrk_traffic.csv | tail -6
