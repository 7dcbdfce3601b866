#!/bin/bash
# round-2 GPU pass 23: backward blend phase 1 software-pipelined; preprocess_bwd at 3 CTAs/SM
mkdir -p gpurun_out
run_bench() {  # name, nvcc extra
  GSB_NVCC_EXTRA="$2" python -m gaussianip_b200.build > /dev/null 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2u_bench_$1.json 2> gpurun_out/r2u_bench_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2u_bench_$1.json").read().strip().splitlines()[-1])
    s=d["roofline"]["stage_us_per_view"]
    print("$1", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "fwd", s["render_fwd"], "bwd", s["render_bwd"], "pbwd", s["preprocess_bwd"])
except Exception as e:
    print("$1 ERR", e); print(open("gpurun_out/r2u_bench_$1.err").read()[-1500:])
PY
}
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_graph.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r2u_tests.txt 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/r2u_tests.txt
run_bench pipe1 ""
run_bench pipe0 "-DGSB_BWD_T2_PIPE1=0"
run_bench pipe1_pbwd3 "-DGSB_PBWD_MINB=3"
run_bench pipe1_again ""
