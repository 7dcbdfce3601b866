#!/bin/bash
# round-2 GPU pass 28: branch-free alpha evaluation (1) and blend (2) in the default forward kernel
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
GSB_NVCC_EXTRA="-DGSB_FWD_NOBRANCH=2" python -m gaussianip_b200.build > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_graph.py tests/test_gpu_renderers.py tests/test_gpu_golden.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r2y_tests_nb2.txt 2>&1
echo "nobranch2 pytest rc $?"; tail -2 gpurun_out/r2y_tests_nb2.txt
run_bench nb2 "-DGSB_FWD_NOBRANCH=2"
run_bench nb1 "-DGSB_FWD_NOBRANCH=1"
run_bench nb2_again "-DGSB_FWD_NOBRANCH=2"
