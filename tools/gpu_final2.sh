#!/bin/bash
# last refresh: whole GPU suite, smoke, LDM bench + launch list
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== LDM bench"; timeout 900 python bench.py --variant ldm --steps 5 --warmup 3 > gpurun_out/bench_ldm.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ldm.log | cut -c1-300
echo "== LDM launch list"
NOPE_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ldm_launches.csv python tools/ldm_time.py 128 tcgen05 > gpurun_out/ldm_ncu.log 2>&1; echo "rc=$?"
