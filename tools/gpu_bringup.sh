#!/bin/bash
# first GPU bring-up: every stage under its own timeout, logs into gpurun_out/
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
python -c "import torch; print(torch.cuda.get_device_name(0))" > gpurun_out/dev.txt 2>&1
echo "== ops (simt conv + memory-bound kernels)"; timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -rA --tb=short > gpurun_out/ops.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/ops.log
echo "== conv tcgen05"; timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -rA --tb=short > gpurun_out/conv_tc.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/conv_tc.log
echo "== unet simt"; NOPE_CONV_IMPL=simt timeout 1200 python -m pytest tests/test_unet_gpu.py -m gpu -q -rA --tb=short > gpurun_out/unet_simt.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/unet_simt.log
echo "== unet tcgen05"; timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -rA --tb=short > gpurun_out/unet_tc.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/unet_tc.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench.log
