#!/bin/bash
# sanitizer logs of the selectable forward variants (pre-cull: racecheck + memcheck; gather4: memcheck)
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
GSB_FWD_VARIANT=precull timeout 120 compute-sanitizer --tool racecheck --log-file gpurun_out/r2_sanitizer_racecheck_precull.log python scripts/sanitize_probe.py > /dev/null 2>&1; grep -E "RACECHECK SUMMARY" gpurun_out/r2_sanitizer_racecheck_precull.log
GSB_FWD_VARIANT=precull timeout 120 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck_precull.log python scripts/sanitize_probe.py > /dev/null 2>&1; grep -E "ERROR SUMMARY" gpurun_out/r2_sanitizer_memcheck_precull.log
GSB_FWD_VARIANT=gather4 timeout 120 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck_gather4.log python scripts/sanitize_probe.py > /dev/null 2>&1; grep -E "ERROR SUMMARY" gpurun_out/r2_sanitizer_memcheck_gather4.log
