#!/bin/bash
# 8 GPUs: parity worker, C3 bench, C4 pipeline (filter + ICE + expected at 3.1 M bins)
N=${1:-8}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29531 tests/multigpu_worker.py small > $O/c8_n${N}_worker_small.log 2>&1
echo "worker small rc=$?" > $O/c8_n${N}_box.log
TB_TIMING=1 timeout 900 $RUN --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > $O/c8_n${N}_bench.json 2> $O/c8_n${N}_bench.err
echo "bench c3 rc=$?" >> $O/c8_n${N}_box.log
TB_TIMING=1 timeout 1500 $RUN --master-port 29534 bench.py --gpus $N --steps 2 --warmup 1 --workload c4 --no-e2e --no-parity > $O/c8_n${N}_bench_c4.json 2> $O/c8_n${N}_bench_c4.err
echo "bench c4 rc=$?" >> $O/c8_n${N}_box.log
grep -h "MULTIGPU_RESULT\|FAIL" $O/c8_n${N}_worker_small.log; cat $O/c8_n${N}_box.log
grep -h "per pass\|bench e2e" $O/c8_n${N}_bench.err | tail -16
grep -h "per pass" $O/c8_n${N}_bench_c4.err | tail -8; tail -3 $O/c8_n${N}_bench_c4.err
python - <<PY
import json
for f in ('$O/c8_n${N}_bench.json','$O/c8_n${N}_bench_c4.json'):
  for line in open(f):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','passes_per_step','converged','final_dev','ms_per_pass','rank_kernel_ms')})
        print(d.get('e2e')); print(d.get('parity')); print(d['pipeline']); print(d['config'])
PY
