#!/bin/bash
# round-2 GPU pass 8: tuning sweep of the transposed backward (slab capacity / occupancy / software pipelining)
mkdir -p gpurun_out
run_bench() {  # name, extra build flags
  GSB_NVCC_EXTRA="$2" python -m gaussianip_b200.build > /dev/null 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2h_bench_$1.json 2> gpurun_out/r2h_bench_$1.err
  echo "$1 rc $?"
}
run_bench c512_b3_p1 ""
run_bench c512_b3_p0 "-DGSB_BWD_T2_PIPE=0"
run_bench c256_b4_p1 "-DGSB_BWD_T2_CAP=256 -DGSB_BWD_T2_MINB=4"
run_bench c256_b4_p0 "-DGSB_BWD_T2_CAP=256 -DGSB_BWD_T2_MINB=4 -DGSB_BWD_T2_PIPE=0"
run_bench c128_b4_p0 "-DGSB_BWD_T2_CAP=128 -DGSB_BWD_T2_MINB=4 -DGSB_BWD_T2_PIPE=0"
python -m gaussianip_b200.build > /dev/null 2>&1
python - <<'PY'
import json
for v in ("c512_b3_p1","c512_b3_p0","c256_b4_p1","c256_b4_p0","c128_b4_p0"):
    try:
        d=json.loads(open(f"gpurun_out/r2h_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["roofline"]["stage_us_per_view"]["render_bwd"])
    except Exception as e:
        print(v, "ERR", e)
PY
