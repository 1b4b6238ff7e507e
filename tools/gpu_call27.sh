#!/bin/bash
# N GPUs, final library: sharded parity workers (small, c3s) incl. K1 one/two streams and N4 from a row-sharded store
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29571 tests/multigpu_worker.py small > $O/c27_n${N}_worker_small.log 2>&1
echo "worker small rc=$?" > $O/c27_n${N}_box.log
timeout 900 $RUN --master-port 29572 tests/multigpu_worker.py c3s > $O/c27_n${N}_worker_c3s.log 2>&1
echo "worker c3s rc=$?" >> $O/c27_n${N}_box.log
grep -h "ok$\|FAIL\|MULTIGPU_RESULT\|exchange in use" $O/c27_n${N}_worker_small.log $O/c27_n${N}_worker_c3s.log; cat $O/c27_n${N}_box.log
free -g | head -2
