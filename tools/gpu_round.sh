#!/bin/bash
# full GPU check: tests, bench, ncu launch list + full captures.  Logs -> gpurun_out/
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -rA --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log
if [ "$1" != "noprof" ]; then
echo "== ncu launch list"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full: conv"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc -c 48 -o gpurun_out/prof_conv -f python tools/profile_step.py > gpurun_out/ncu_conv.log 2>&1; echo "rc=$?"
echo "== ncu full: memory-bound kernels"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'gn_|linattn|final_conv|upsample|bcast|midattn' -c 40 -o gpurun_out/prof_mem -f python tools/profile_step.py > gpurun_out/ncu_mem.log 2>&1; echo "rc=$?"
ls -la gpurun_out
fi
