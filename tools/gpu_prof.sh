#!/bin/bash
# tests + bench + compact ncu exports (csv only; reports stay on the box)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log
for c in $BENCH_CHUNKS; do
  echo "== bench chunk $c"; timeout 600 python bench.py --steps 5 --warmup 3 --chunk $c --no-cpu-baseline > gpurun_out/bench_chunk$c.log 2>&1; tail -1 gpurun_out/bench_chunk$c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['conv_ms_per_step'], d['e2e']['value'])"
done
echo "== ncu launch list"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full: conv"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc -c 40 -o /tmp/prof_conv -f python tools/profile_step.py > gpurun_out/ncu_conv.log 2>&1; echo "rc=$?"
ncu -i /tmp/prof_conv.ncu-rep --page raw --csv > gpurun_out/prof_conv_raw.csv 2>/dev/null
ncu -i /tmp/prof_conv.ncu-rep --page source --csv --kernel-id :::3 > gpurun_out/prof_conv_source_k3.csv 2>/dev/null
echo "== ncu full: memory-bound kernels"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'gn_|linattn|final_conv|bcast' -c 24 -o /tmp/prof_mem -f python tools/profile_step.py > gpurun_out/ncu_mem.log 2>&1; echo "rc=$?"
ncu -i /tmp/prof_mem.ncu-rep --page raw --csv > gpurun_out/prof_mem_raw.csv 2>/dev/null
ncu -i /tmp/prof_mem.ncu-rep --page source --csv --kernel-id :::2 > gpurun_out/prof_mem_source_k2.csv 2>/dev/null
du -sh gpurun_out
