#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== compute-sanitizer memcheck (smoke: 6-pose predict_pose incl. encoder)"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 8 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|smoke:|at nope" gpurun_out/sanitizer_memcheck.log | sort | uniq -c | head
echo "== test_shapeNet.py (synthetic)"; timeout 600 python test_shapeNet.py --batches 2 --batch-size 2 --grid 642 > gpurun_out/test_shapenet.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/test_shapenet.log | cut -c1-500
echo "== unet + ops tests"; timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_ops_gpu.py -m gpu -q --tb=short > gpurun_out/t.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t.log
