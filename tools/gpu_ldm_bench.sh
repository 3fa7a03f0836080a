#!/bin/bash
# LDM variant: bench line + ncu full captures (attention, GEMMs), csv exports only
mkdir -p gpurun_out
timeout 900 python bench.py --variant ldm --steps 5 --warmup 3 > gpurun_out/bench_ldm.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_ldm.log | cut -c1-1500
NOPE_PROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ldm_attn_tc -c 6 -o /tmp/ldm_attn -f python tools/ldm_time.py 128 tcgen05 > gpurun_out/ldm_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ncu -i /tmp/ldm_attn.ncu-rep --page raw --csv > gpurun_out/ldm_attn_raw.csv 2>/dev/null
ncu -i /tmp/ldm_attn.ncu-rep --page source --csv --kernel-id :::2 > gpurun_out/ldm_attn_source_k2.csv 2>/dev/null
NOPE_PROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:conv_tc2 -c 60 -o /tmp/ldm_conv -f python tools/ldm_time.py 128 tcgen05 > gpurun_out/ldm_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
ncu -i /tmp/ldm_conv.ncu-rep --page raw --csv > gpurun_out/ldm_conv_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out | head -40
