#!/bin/bash
# N GPUs: bench of the final library of round 2 (K1 on two streams)
N=${1:-4}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29581 bench.py --gpus $N --steps 3 --warmup 3 > $O/c34_n${N}_bench.json 2> $O/c34_n${N}_bench.err
echo "bench rc=$?" > $O/c34_n${N}_box.log
cat $O/c34_n${N}_box.log
python - <<PY
import json
for line in open('$O/c34_n${N}_bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print({k:d[k] for k in ('value','ms_per_step','passes_per_step','converged','ms_per_pass','rank_kernel_ms')}, (d.get('parity') or {}).get('ok'), (d.get('e2e') or {}))
PY
tail -3 $O/c34_n${N}_bench.err
