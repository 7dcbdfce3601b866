#!/bin/bash
# round-2 GPU pass 22: forward blend with TMA gather4 row gathers vs cp.async, at 3 and 4 CTAs per SM
mkdir -p gpurun_out
run_bench() {  # name, nvcc extra, fwd variant
  GSB_NVCC_EXTRA="$2" python -m gaussianip_b200.build > /dev/null 2>&1
  GSB_FWD_VARIANT=$3 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2t_bench_$1.json 2> gpurun_out/r2t_bench_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2t_bench_$1.json").read().strip().splitlines()[-1])
    print("$1", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "fwd", d["roofline"]["stage_us_per_view"]["render_fwd"], "bwd", d["roofline"]["stage_us_per_view"]["render_bwd"])
except Exception as e:
    print("$1 ERR", e); print(open("gpurun_out/r2t_bench_$1.err").read()[-1500:])
PY
}
python -m gaussianip_b200.build > /dev/null 2>&1
GSB_FWD_VARIANT=gather4 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_graph.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r2t_tests_gather4.txt 2>&1
echo "gather4 pytest rc $?"; tail -3 gpurun_out/r2t_tests_gather4.txt
run_bench cpasync_3cta "" per_hit
run_bench gather4_3cta "" gather4
run_bench cpasync_4cta "-DGSB_FWD_MINB=4" per_hit
run_bench gather4_4cta "-DGSB_FWD_MINB=4" gather4
