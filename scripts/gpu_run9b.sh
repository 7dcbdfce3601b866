#!/bin/bash
# compute-sanitizer logs of the probe workload (memcheck, racecheck, initcheck)
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r2_sanitizer_$tool.log python scripts/sanitize_probe.py > gpurun_out/r2_sanitizer_$tool.out 2>&1
  echo "$tool rc $?"; tail -2 gpurun_out/r2_sanitizer_$tool.out; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2_sanitizer_$tool.log
done
