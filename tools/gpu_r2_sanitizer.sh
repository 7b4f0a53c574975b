#!/bin/bash
# compute-sanitizer over one small call through every kernel family (tools/sanitizer_subset.py); $1 = memcheck | racecheck | synccheck | initcheck
tool=${1:-memcheck}
mkdir -p gpurun_out
python tools/sanitizer_subset.py $2 > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
tail -1 gpurun_out/sanitizer_plain.log | cut -c1-300
t0=$(date +%s)
timeout ${SAN_TIMEOUT:-1500} compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 20 python tools/sanitizer_subset.py $2 > gpurun_out/r2_$tool.log 2>&1
rc=$?
echo "$tool exit $rc after $(( $(date +%s) - t0 )) s"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer_subset" gpurun_out/r2_$tool.log | cut -c1-400
grep -E "=========" gpurun_out/r2_$tool.log | head -40 | cut -c1-220
