#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== racecheck (smoke)"; timeout 1500 compute-sanitizer --tool racecheck --print-limit 6 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|smoke:|hazard|at nope" gpurun_out/sanitizer_racecheck.log | sort | uniq -c | sort -rn | head -12
echo "== synccheck (smoke)"; timeout 1500 compute-sanitizer --tool synccheck --print-limit 6 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_synccheck.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|smoke:|at nope|Barrier" gpurun_out/sanitizer_synccheck.log | sort | uniq -c | sort -rn | head -8
