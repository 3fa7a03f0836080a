#!/bin/bash
# source-level ncu captures of three LDM launches: GEGLU GEMM at 32^2, a residual 1x1 GEMM at 32^2, attention at 32^2
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip
  NOPE_PROFILE=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/src_$1 -f python tools/ldm_time.py 128 tcgen05 > gpurun_out/ldm_src_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i /tmp/src_$1.ncu-rep --page source --csv > gpurun_out/ldm_src_$1_source.csv 2>/dev/null
  ncu -i /tmp/src_$1.ncu-rep --page raw --csv > gpurun_out/ldm_src_$1_raw.csv 2>/dev/null
}
cap geglu conv_tc2 5
cap resid conv_tc2 12
cap attn ldm_attn_tc 1
du -sh gpurun_out; ls -la gpurun_out | grep ldm_src
