#!/bin/bash
# LL exchange: tests on one GPU of the box, parity workers and bench at N, flags protocol for comparison
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests -m gpu -q --maxfail=8 -k "not full_size and not batch_of_64 and not coverage" > $O/c18_tests.log 2>&1
echo "tests rc=$?" > $O/c18_n${N}_box.log
timeout 600 $RUN --master-port 29551 tests/multigpu_worker.py small > $O/c18_n${N}_worker_small.log 2>&1
echo "worker small rc=$?" >> $O/c18_n${N}_box.log
timeout 900 $RUN --master-port 29552 tests/multigpu_worker.py c3s > $O/c18_n${N}_worker_c3s.log 2>&1
echo "worker c3s rc=$?" >> $O/c18_n${N}_box.log
TB_TIMING=1 timeout 900 $RUN --master-port 29553 bench.py --gpus $N --steps 3 --warmup 2 > $O/c18_n${N}_bench.json 2> $O/c18_n${N}_bench.err
echo "bench rc=$?" >> $O/c18_n${N}_box.log
TB_EXCHANGE=flags TB_TIMING=1 timeout 900 $RUN --master-port 29554 bench.py --gpus $N --steps 3 --warmup 2 --no-parity --no-e2e > $O/c18_n${N}_bench_flags.json 2> $O/c18_n${N}_bench_flags.err
echo "bench flags rc=$?" >> $O/c18_n${N}_box.log
tail -2 $O/c18_tests.log; grep -h "MULTIGPU_RESULT\|FAIL" $O/c18_n${N}_worker_small.log $O/c18_n${N}_worker_c3s.log; cat $O/c18_n${N}_box.log
echo "-- LL"; grep -h "per pass\|CTA 0" $O/c18_n${N}_bench.err | tail -4
echo "-- flags"; grep -h "per pass\|CTA 0" $O/c18_n${N}_bench_flags.err | tail -4
python - <<PY
import json, os
for f in ('$O/c18_n${N}_bench.json','$O/c18_n${N}_bench_flags.json'):
  for line in open(f):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','passes_per_step','converged','ms_per_pass')}, (d.get('parity') or {}).get('ok'), (d.get('e2e') or {}).get('s_per_step'))
PY
