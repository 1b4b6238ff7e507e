#!/bin/bash
# 2 GPUs: sharded parity (incl. K1 one/two streams, N4 from a row-sharded store), bench at N=2 with the two-stream K1
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29561 tests/multigpu_worker.py small > $O/c23_n${N}_worker_small.log 2>&1
echo "worker small rc=$?" > $O/c23_n${N}_box.log
timeout 900 $RUN --master-port 29562 tests/multigpu_worker.py c3s > $O/c23_n${N}_worker_c3s.log 2>&1
echo "worker c3s rc=$?" >> $O/c23_n${N}_box.log
timeout 900 $RUN --master-port 29563 bench.py --gpus $N --steps 3 --warmup 3 > $O/c23_n${N}_bench.json 2> $O/c23_n${N}_bench.err
echo "bench rc=$?" >> $O/c23_n${N}_box.log
grep -h "ok$\|FAIL\|MULTIGPU_RESULT\|exchange in use" $O/c23_n${N}_worker_small.log $O/c23_n${N}_worker_c3s.log; cat $O/c23_n${N}_box.log
tail -5 $O/c23_n${N}_worker_small.log
python - <<PY
import json
for line in open('$O/c23_n${N}_bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','passes_per_step','converged','ms_per_pass','rank_kernel_ms')}, (d.get('parity') or {}), (d.get('e2e') or {}))
PY
tail -3 $O/c23_n${N}_bench.err
