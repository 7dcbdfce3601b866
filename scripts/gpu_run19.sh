#!/bin/bash
# diagnostic: phase timing of one radix pass + sort bench only
mkdir -p gpurun_out
GSB_NVCC_EXTRA="-DGSB_RADIX_TIMING $GSB_EXTRA" python -m gaussianip_b200.build > /dev/null 2>&1
timeout 120 python scripts/radix_timing.py 1000000 32 2>&1 | head -9
GSB_NVCC_EXTRA="$GSB_EXTRA" python -m gaussianip_b200.build > /dev/null 2>&1
timeout 300 python scripts/sort_bench.py 2>&1 | grep "n=  1048576\|n=  2097152"
