#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== width tests"; timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q --tb=short -k "widths" > gpurun_out/t_w.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_w.log
