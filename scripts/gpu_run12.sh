#!/bin/bash
# round-2 GPU pass 12: look-back width of the onesweep radix passes
mkdir -p gpurun_out
run_bench() {  # name, extra build flags
  GSB_NVCC_EXTRA="$2" python -m gaussianip_b200.build > /dev/null 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2l_bench_$1.json 2> gpurun_out/r2l_bench_$1.err
  echo "$1 rc $?"
}
run_bench lb4 ""
run_bench lb8 "-DGSB_RADIX_LB=8"
run_bench lb16 "-DGSB_RADIX_LB=16"
run_bench lb32 "-DGSB_RADIX_LB=32"
GSB_NVCC_EXTRA="-DGSB_RADIX_LB=16" python -m gaussianip_b200.build > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_knn.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "forward_exact or knn" > gpurun_out/r2l_tests.txt 2>&1
tail -3 gpurun_out/r2l_tests.txt
python -m gaussianip_b200.build > /dev/null 2>&1
python - <<'PY'
import json
for v in ("lb4","lb8","lb16","lb32"):
    try:
        d=json.loads(open(f"gpurun_out/r2l_bench_{v}.json").read().strip().splitlines()[-1])
        st=d["roofline"]["stage_us_per_view"]
        print(v, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k: st[k] for k in ("depth_sort","scan_emit","tile_sort","ranges")})
    except Exception as e:
        print(v, "ERR", e)
PY
