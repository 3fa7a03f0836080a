#!/bin/bash
mkdir -p gpurun_out
timeout 120 python bench.py --variant ldm --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/bench_ldm_n1_check.log 2>&1; echo "n1 rc=$?"; tail -1 gpurun_out/bench_ldm_n1_check.log | cut -c1-200
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --variant ldm --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_ldm_n2.log 2>&1; echo "n2 rc=$?"; tail -1 gpurun_out/bench_ldm_n2.log | cut -c1-600
