#!/bin/bash
# round-2 GPU pass 34 (last): forward at 4 CTAs/SM with the branch-free loop; radix polling without nanosleep
mkdir -p gpurun_out
run_bench() {  # name, nvcc extra
  GSB_NVCC_EXTRA="$2" python -m gaussianip_b200.build > /dev/null 2>&1
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2y_bench_$1.json 2> gpurun_out/r2y_bench_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2y_bench_$1.json").read().strip().splitlines()[-1])
    s=d["roofline"]["stage_us_per_view"]
    print("$1", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "fwd", s["render_fwd"], "depth", s["depth_sort"], "tile", s["tile_sort"])
except Exception as e:
    print("$1 ERR", e)
PY
}
run_bench fwd4cta "-DGSB_FWD_MINB=4"
run_bench poll0 "-DGSB_RADIX_POLL_NS=0"
