#!/bin/bash
# round-2 GPU pass 31: warp-uniform phase 2 in the transposed backward
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
GSB_NVCC_EXTRA="-DGSB_BWD_P2_UNIFORM=1" python -m gaussianip_b200.build > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "backward or config or batched or fused or nan or linearity" > gpurun_out/r2y_tests_p2u.txt 2>&1
echo "p2 uniform pytest rc $?"; tail -2 gpurun_out/r2y_tests_p2u.txt
run_bench p2u "-DGSB_BWD_P2_UNIFORM=1"
run_bench p2d ""
