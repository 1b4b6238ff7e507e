#!/bin/bash
# N GPUs: quick parity worker + bench (fused exchange), TB_TIMING split
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29521 tests/multigpu_worker.py small > $O/c6_n${N}_worker_small.log 2>&1
echo "worker small rc=$?" > $O/c6_n${N}_box.log
TB_TIMING=1 timeout 900 $RUN --master-port 29523 bench.py --gpus $N --steps 3 --warmup 2 > $O/c6_n${N}_bench.json 2> $O/c6_n${N}_bench.err
echo "bench rc=$?" >> $O/c6_n${N}_box.log
grep -h "MULTIGPU_RESULT\|FAIL" $O/c6_n${N}_worker_small.log; cat $O/c6_n${N}_box.log
grep -h "per pass\|bench e2e" $O/c6_n${N}_bench.err | tail -8
python - <<PY
import json
for line in open('$O/c6_n${N}_bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','passes_per_step','converged','ms_per_pass','rank_kernel_ms')})
        print(d['e2e']); print(d['parity'])
PY
