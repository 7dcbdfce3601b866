#!/bin/bash
# diagnostic: phase timing of one radix pass, then the regular pass
mkdir -p gpurun_out
GSB_NVCC_EXTRA="-DGSB_RADIX_TIMING" python -m gaussianip_b200.build > /dev/null 2>&1
timeout 120 python scripts/radix_timing.py 1000000 32 > gpurun_out/r2q_radix_timing.txt 2>&1
timeout 120 python scripts/radix_timing.py 2000000 12 >> gpurun_out/r2q_radix_timing.txt 2>&1
cat gpurun_out/r2q_radix_timing.txt
bash scripts/gpu_run17.sh
