#!/bin/bash
# LDM variant: parity tests + timing (+ launch list)
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
timeout 1200 python -m pytest tests/test_ldm_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/ldm_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/ldm_tests.log
tail -40 gpurun_out/ldm_tests.log
cp gpurun_out/parity_log.jsonl gpurun_out/ldm_parity_log.jsonl 2>/dev/null
timeout 300 python tools/ldm_time.py 128 tcgen05 > gpurun_out/ldm_time.log 2>&1
timeout 300 python tools/ldm_time.py 512 tcgen05 >> gpurun_out/ldm_time.log 2>&1
cat gpurun_out/ldm_time.log
NOPE_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ldm_launches.csv python tools/ldm_time.py 128 tcgen05 > gpurun_out/ldm_ncu.log 2>&1
echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/ldm_launches.csv | head -40
