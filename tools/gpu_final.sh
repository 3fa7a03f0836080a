#!/bin/bash
# final evidence refresh: full GPU tests, smoke, bench (+reference arm), launch list, conv ncu
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
echo "== ncu launch list"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== LDM bench"; timeout 900 python bench.py --variant ldm --steps 5 --warmup 3 > gpurun_out/bench_ldm.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ldm.log | cut -c1-400
echo "== LDM launch list"
NOPE_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ldm_launches.csv python tools/ldm_time.py 128 tcgen05 > gpurun_out/ldm_ncu.log 2>&1; echo "rc=$?"
du -sh gpurun_out
echo "== compute-sanitizer memcheck: LDM tests"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_ldm_gpu.py -q -m gpu -x -p no:cacheprovider > gpurun_out/sanitizer_memcheck_ldm.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_ldm.log
