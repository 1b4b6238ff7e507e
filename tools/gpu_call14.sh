#!/bin/bash
# final multi-GPU session: parity worker, C3 bench, optional C4 pipeline
N=${1:-8}; C4=${2:-yes}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29541 tests/multigpu_worker.py small > $O/c14_n${N}_worker_small.log 2>&1
echo "worker small rc=$?" > $O/c14_n${N}_box.log
timeout 900 $RUN --master-port 29542 tests/multigpu_worker.py c3s > $O/c14_n${N}_worker_c3s.log 2>&1
echo "worker c3s rc=$?" >> $O/c14_n${N}_box.log
TB_TIMING=1 timeout 900 $RUN --master-port 29543 bench.py --gpus $N --steps 5 --warmup 3 > $O/c14_n${N}_bench.json 2> $O/c14_n${N}_bench.err
echo "bench c3 rc=$?" >> $O/c14_n${N}_box.log
if [ "$C4" = yes ]; then
  TB_TIMING=1 timeout 1500 $RUN --master-port 29544 bench.py --gpus $N --steps 2 --warmup 1 --workload c4 --no-e2e --no-parity > $O/c14_n${N}_bench_c4.json 2> $O/c14_n${N}_bench_c4.err
  echo "bench c4 rc=$?" >> $O/c14_n${N}_box.log
fi
grep -h "MULTIGPU_RESULT\|FAIL" $O/c14_n${N}_worker_small.log $O/c14_n${N}_worker_c3s.log; cat $O/c14_n${N}_box.log
grep -h "per pass\|CTA 0" $O/c14_n${N}_bench.err | tail -4
[ "$C4" = yes ] && { grep -h "per pass" $O/c14_n${N}_bench_c4.err | tail -2; tail -2 $O/c14_n${N}_bench_c4.err; }
python - <<PY
import json, os
for f in ('$O/c14_n${N}_bench.json','$O/c14_n${N}_bench_c4.json'):
  if not os.path.exists(f): continue
  for line in open(f):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','passes_per_step','converged','final_dev','ms_per_pass','rank_kernel_ms')})
        print(d.get('e2e')); print((d.get('parity') or {}).get('ok')); print(d['pipeline']); print(d['config'])
PY
