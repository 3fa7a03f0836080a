#!/bin/bash
# LDM variant: timing + ncu launch list
mkdir -p gpurun_out
timeout 300 python tools/ldm_time.py 128 tcgen05 > gpurun_out/ldm_time.log 2>&1
timeout 300 python tools/ldm_time.py 256 tcgen05 >> gpurun_out/ldm_time.log 2>&1
cat gpurun_out/ldm_time.log
NOPE_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ldm_launches.csv python tools/ldm_time.py 128 tcgen05 > gpurun_out/ldm_ncu.log 2>&1
echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/ldm_launches.csv | head -40
