#!/bin/bash
# multi-GPU check: NCCL sharded sweep correctness + bench at N GPUs.  usage: gpu_multi.sh N
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== dist check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "rc=$?"; grep -E "DIST_CHECK|Error|error" gpurun_out/dist_check_$N.log | head -5
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n$N.log
