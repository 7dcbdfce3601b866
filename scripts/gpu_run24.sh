#!/bin/bash
# round-2 GPU pass 24: forward blend behind a CTA-wide pre-cull
mkdir -p gpurun_out
run_bench() {  # name, nvcc extra, fwd variant
  GSB_NVCC_EXTRA="$2" python -m gaussianip_b200.build > /dev/null 2>&1
  GSB_FWD_VARIANT=$3 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $4 > gpurun_out/r2v_bench_$1.json 2> gpurun_out/r2v_bench_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2v_bench_$1.json").read().strip().splitlines()[-1])
    s=d["roofline"]["stage_us_per_view"]
    print("$1", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "fwd", s["render_fwd"], "bwd", s["render_bwd"], "vcr", round((d.get("vcr") or {}).get("value",0),1))
except Exception as e:
    print("$1 ERR", e); print(open("gpurun_out/r2v_bench_$1.err").read()[-1500:])
PY
}
python -m gaussianip_b200.build > /dev/null 2>&1
GSB_FWD_VARIANT=precull timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_graph.py tests/test_gpu_renderers.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r2v_tests_precull.txt 2>&1
echo "precull pytest rc $?"; tail -4 gpurun_out/r2v_tests_precull.txt
run_bench per_hit "" per_hit --no-vcr
run_bench precull "" precull
run_bench precull_4cta "-DGSB_FWD_MINB=4" precull --no-vcr
