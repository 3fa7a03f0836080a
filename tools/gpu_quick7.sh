#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
echo "== unet tests"; timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_conv_tc2_gpu.py -m gpu -q --tb=short > gpurun_out/t.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('hyp/s', d['value'], 'ms', d['ms_per_step'], 'conv TF/s', d['roofline']['achieved'], 'conv ms', d['roofline']['conv_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
echo "== ncu launch list"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
