#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
echo "== unet + ops + conv tests"; timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_ops_gpu.py tests/test_conv_tc2_gpu.py -m gpu -q --tb=short > gpurun_out/t.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t.log
grep -E "cfg1_golden|level2_642|grid26" gpurun_out/parity_log.jsonl | cut -c1-300
echo "== unet tests, 1-CTA kernel"; NOPE_CONV_IMPL=tcgen05 timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q --tb=short > gpurun_out/t1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/t1.log
echo "== unet tests, simt twin"; NOPE_CONV_IMPL=simt timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q --tb=short -k "cfg1 or taps or grid26" > gpurun_out/t2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/t2.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('hyp/s', d['value'], 'ms', d['ms_per_step'], 'conv TF/s', d['roofline']['achieved'], 'conv ms', d['roofline']['conv_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
