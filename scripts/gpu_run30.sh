#!/bin/bash
# round-2 GPU pass 30: radix ranking without the leader branch / shuffle
mkdir -p gpurun_out
for nb in 1 0; do
  GSB_NVCC_EXTRA="-DGSB_RADIX_RANK_NOBRANCH=$nb" python -m gaussianip_b200.build > /dev/null 2>&1
  echo "rank_nobranch=$nb"; timeout 300 python scripts/sort_bench.py 2>&1 | grep "n=  1048576\|n=  2097152\|correct=False"
done
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2y_bench_ranknb.json 2> gpurun_out/r2y_bench_ranknb.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2y_bench_ranknb.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), d["roofline"]["stage_us_per_view"])
PY
