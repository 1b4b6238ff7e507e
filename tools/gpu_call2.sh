#!/bin/bash
# round-2 GPU call (N GPUs): sharded parity worker + bench at N
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L > $O/c2_n${N}_box.log 2>&1
nvidia-smi topo -m >> $O/c2_n${N}_box.log 2>&1
timeout 600 $RUN --master-port 29511 tests/multigpu_worker.py small > $O/c2_n${N}_worker_small.log 2>&1
echo "worker small rc=$?" >> $O/c2_n${N}_box.log
timeout 900 $RUN --master-port 29512 tests/multigpu_worker.py c3s > $O/c2_n${N}_worker_c3s.log 2>&1
echo "worker c3s rc=$?" >> $O/c2_n${N}_box.log
TB_TIMING=1 timeout 900 $RUN --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 > $O/c2_n${N}_bench.json 2> $O/c2_n${N}_bench.err
echo "bench rc=$?" >> $O/c2_n${N}_box.log
TB_EXCHANGE=nccl TB_TIMING=1 timeout 900 $RUN --master-port 29514 bench.py --gpus $N --steps 3 --warmup 2 --no-parity --no-e2e > $O/c2_n${N}_bench_nccl.json 2> $O/c2_n${N}_bench_nccl.err
echo "bench nccl rc=$?" >> $O/c2_n${N}_box.log
grep -h "MULTIGPU_RESULT\|FAIL\|exchange in use" $O/c2_n${N}_worker_small.log $O/c2_n${N}_worker_c3s.log; cat $O/c2_n${N}_box.log | tail -6
head -c 2500 $O/c2_n${N}_bench.json; echo; grep -h "per pass\|bench e2e\|Error\|error" $O/c2_n${N}_bench.err | tail -12; grep -h "per pass" $O/c2_n${N}_bench_nccl.err | tail -4
