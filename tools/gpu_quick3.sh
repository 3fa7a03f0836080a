#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
echo "== new tests"; timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q --tb=short -k "level2 or native_encoder or shard" > gpurun_out/t_new.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_new.log; grep level2_642 gpurun_out/parity_log.jsonl | cut -c1-600
echo "== compute-sanitizer memcheck (smoke: 6-pose predict_pose incl. encoder)"
timeout 1200 compute-sanitizer --tool memcheck --print-limit 8 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
echo "== bench (with eager GPU baseline)"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('hyp/s', d['value'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'], 'eager', d['eager_gpu_baseline'])"
echo "== bench B=8 N=2562 (configs[2] size, fp16)"; timeout 900 python bench.py --steps 3 --warmup 3 --queries 8 --poses 2562 --no-cpu-baseline > gpurun_out/bench_b8_n2562.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_b8_n2562.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('hyp/s', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'conv TF/s', d['roofline']['achieved'])"
