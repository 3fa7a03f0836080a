#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== width tests"; timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q --tb=short -k "widths" > gpurun_out/t_w.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_w.log
timeout 300 python tools/debug_width.py 64 simt final_res_block,final_conv.0 2>&1 | tail -3
timeout 300 python tools/debug_width.py 128 tcgen05 downs.0.2,downs.0.2,final_conv.0 2>&1 | tail -4
echo "== all unet tests"; timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q --tb=short > gpurun_out/t_u.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_u.log
