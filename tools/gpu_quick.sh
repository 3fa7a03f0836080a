#!/bin/bash
# quick GPU check of a kernel change: conv tests, unet tests, bench (+ optional 2-CTA leg)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
echo "== conv + ops tests"; timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_ops_gpu.py -m gpu -q --tb=short -x > gpurun_out/t_conv.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_conv.log
echo "== unet tests"; timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q --tb=short > gpurun_out/t_unet.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_unet.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('hyp/s', d['value'], 'ms', d['ms_per_step'], 'conv TF/s', d['roofline']['achieved'], 'conv ms', d['roofline']['conv_ms_per_step'], 'e2e', d['e2e']['value'])"
echo "== 2-CTA tests"; timeout 600 python -m pytest tests/test_conv_tc2_gpu.py -m gpu -q --tb=short > gpurun_out/t_conv2.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_conv2.log
echo "== unet tests 1cta"; NOPE_CONV_IMPL=tcgen05 timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q --tb=short > gpurun_out/t_unet2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_unet2.log
echo "== bench 1cta"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-impl tcgen05 > gpurun_out/bench_1cta.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_1cta.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('hyp/s', d['value'], 'ms', d['ms_per_step'], 'conv TF/s', d['roofline']['achieved'], 'conv ms', d['roofline']['conv_ms_per_step'], 'e2e', d['e2e']['value'])"
echo "== ncu launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
