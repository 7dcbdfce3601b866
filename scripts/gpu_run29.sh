#!/bin/bash
# round-2 GPU pass 29: branch-free forward (default now) + branch-free phase 1 of the transposed backward
mkdir -p gpurun_out
run_bench() {  # name, nvcc extra
  GSB_NVCC_EXTRA="$2" python -m gaussianip_b200.build > /dev/null 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2y_bench_$1.json 2> gpurun_out/r2y_bench_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2y_bench_$1.json").read().strip().splitlines()[-1])
    s=d["roofline"]["stage_us_per_view"]
    print("$1", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "fwd", s["render_fwd"], "bwd", s["render_bwd"])
except Exception as e:
    print("$1 ERR", e); print(open("gpurun_out/r2y_bench_$1.err").read()[-1500:])
PY
}
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r2y_tests_all.txt 2>&1
echo "pytest rc $?"; tail -2 gpurun_out/r2y_tests_all.txt
run_bench bwd_nb1 ""
run_bench bwd_nb0 "-DGSB_BWD_NOBRANCH=0"
