#!/bin/bash
# tests + bench + compact ncu exports (csv only; big reports stay on the box)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-400
echo "== ncu launch list"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full: conv (sweep)"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc -s 56 -c 82 -o /tmp/prof_conv -f python tools/profile_step.py > gpurun_out/ncu_conv.log 2>&1; echo "rc=$?"
ncu -i /tmp/prof_conv.ncu-rep --page raw --csv > gpurun_out/prof_conv_raw.csv 2>/dev/null
ncu -i /tmp/prof_conv.ncu-rep --page source --csv --kernel-id :::3 > gpurun_out/prof_conv_source_k3.csv 2>/dev/null
echo "== ncu full: memory-bound kernels + encoder convs"
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'gn_|linattn|final_conv|bcast|enc_' -c 30 -o /tmp/prof_mem -f python tools/profile_step.py > gpurun_out/ncu_mem.log 2>&1; echo "rc=$?"
ncu -i /tmp/prof_mem.ncu-rep --page raw --csv > gpurun_out/prof_mem_raw.csv 2>/dev/null
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:conv_tc -c 56 -o /tmp/prof_enc -f python tools/profile_step.py > gpurun_out/ncu_enc.log 2>&1; echo "rc=$?"
ncu -i /tmp/prof_enc.ncu-rep --page raw --csv > gpurun_out/prof_enc_raw.csv 2>/dev/null
# one small report kept as .ncu-rep for the record: 3 launches of the dominant conv shape
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc -s 58 -c 3 -o gpurun_out/prof_conv_3k -f python tools/profile_step.py > /dev/null 2>&1
du -sh gpurun_out
