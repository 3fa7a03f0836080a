#!/bin/bash
# whole GPU suite + LDM timing + launch list
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/ldm_time.py 128 tcgen05 > gpurun_out/ldm_time.log 2>&1
timeout 300 python tools/ldm_time.py 512 tcgen05 >> gpurun_out/ldm_time.log 2>&1
cat gpurun_out/ldm_time.log
NOPE_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ldm_launches.csv python tools/ldm_time.py 128 tcgen05 > gpurun_out/ldm_ncu.log 2>&1
echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/ldm_launches.csv | head -24
