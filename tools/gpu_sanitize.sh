#!/bin/bash
# whole GPU test-suite under compute-sanitizer memcheck (slow: ~10x)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 9 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/sanitizer_suite.log 2>&1; echo "rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|at nope" gpurun_out/sanitizer_suite.log | sort | uniq -c | sort -rn | head -12
