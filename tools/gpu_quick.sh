#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 120 python tools/sanitize_updates.py > gpurun_out/sanitize_updates_$tool.log 2>&1
echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^pr|^ct|^radon|hazard|Invalid" gpurun_out/sanitize_updates_$tool.log | sort | uniq -c | head -12
done
