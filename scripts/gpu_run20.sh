#!/bin/bash
# round-2 GPU pass 20: persistent emit / preprocess grids, rect-derived instance counts
mkdir -p gpurun_out
bash scripts/gpu_run17.sh
for mb in 5; do
  GSB_NVCC_EXTRA="-DGSB_EMIT_MINB=$mb" python -m gaussianip_b200.build > /dev/null 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2r_bench_mb$mb.json 2> gpurun_out/r2r_bench_mb$mb.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2r_bench_mb$mb.json").read().strip().splitlines()[-1])
print("MINB $mb value", round(d["value"],1), d["roofline"]["stage_us_per_view"])
PY
done
