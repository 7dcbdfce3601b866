#!/bin/bash
# round-2 GPU pass 15: occupancy of the forward blend (registers capped for 5 / 6 CTAs per SM)
mkdir -p gpurun_out
run_bench() {  # name, extra build flags
  GSB_NVCC_EXTRA="$2" python -m gaussianip_b200.build > /dev/null 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2n_bench_$1.json 2> gpurun_out/r2n_bench_$1.err
  echo "$1 rc $?"
}
run_bench b4 ""
run_bench b5 "-DGSB_FWD_MINB=5"
run_bench b6 "-DGSB_FWD_MINB=6"
python -m gaussianip_b200.build > /dev/null 2>&1
python - <<'PY'
import json
for v in ("b4","b5","b6"):
    try:
        d=json.loads(open(f"gpurun_out/r2n_bench_{v}.json").read().strip().splitlines()[-1])
        st=d["roofline"]["stage_us_per_view"]
        print(v, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), st["render_fwd"], st["render_bwd"])
    except Exception as e:
        print(v, "ERR", e)
PY
