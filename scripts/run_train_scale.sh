#!/bin/bash
# usage: scripts/run_train_scale.sh N "configs..." : DDP training-step bench at N GPUs, JSON lines + NCCL log tail into gpurun_out/
N=$1; CFGS=${2:-"3 2"}
mkdir -p gpurun_out
for c in $CFGS; do
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
     scripts/train_step_bench.py --config $c --steps 8 --warmup 3 > gpurun_out/train_c${c}_n${N}.log 2>&1
  grep '^{"metric"' gpurun_out/train_c${c}_n${N}.log | tee gpurun_out/train_c${c}_n${N}.json
  grep -E "NCCL INFO (AllReduce|Connected|comm|Channel|NVLS|ncclCommInitRank)" gpurun_out/train_c${c}_n${N}.log | grep -v "Channel [0-9][0-9]/" | tail -12 > gpurun_out/train_c${c}_n${N}_nccl_tail.txt
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 \
     scripts/train_step_bench.py --config $c --steps 8 --warmup 3 --no-allreduce 2>/dev/null | grep '^{"metric"' > gpurun_out/train_c${c}_n${N}_noar.json
  python - <<PY
import json
a=json.load(open("gpurun_out/train_c${c}_n${N}.json")); b=json.load(open("gpurun_out/train_c${c}_n${N}_noar.json"))
print("config $c N=$N: %.2f ms/step with all-reduce, %.2f ms without (replicas drift), exposed wait %.2f ms, in sync %s" % (a["ms_per_step"], b["ms_per_step"], a["allreduce"]["exposed_ms_per_step"], a["replicas_in_sync"]))
PY
done
tail -5 gpurun_out/train_c3_n${N}_nccl_tail.txt
