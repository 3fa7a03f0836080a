#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
echo "== encoder test"; timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q --tb=short -k "native_encoder" > gpurun_out/t_enc.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_enc.log; grep native_encoder gpurun_out/parity_log.jsonl
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_all.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('hyp/s', d['value'], 'ms', d['ms_per_step'], 'conv TF/s', d['roofline']['achieved'], 'conv ms', d['roofline']['conv_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
