#!/bin/bash
# round-2 GPU pass 26: pre-culled forward blend with ids staged in shared memory, 2 or 4 hit streams per warp
mkdir -p gpurun_out
for u in 2 4; do
  GSB_NVCC_EXTRA="-DGSB_PC_UNITS=$u" python -m gaussianip_b200.build > /dev/null 2>&1
  GSB_FWD_VARIANT=precull timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_graph.py tests/test_gpu_renderers.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r2w_tests_precull_u$u.txt 2>&1
  echo "units $u pytest rc $?"; tail -2 gpurun_out/r2w_tests_precull_u$u.txt
  GSB_FWD_VARIANT=precull timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2w_bench_u$u.json 2> gpurun_out/r2w_bench_u$u.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2w_bench_u$u.json").read().strip().splitlines()[-1])
    s=d["roofline"]["stage_us_per_view"]
    print("units $u", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "fwd", s["render_fwd"], "bwd", s["render_bwd"])
except Exception as e:
    print("units $u ERR", e); print(open("gpurun_out/r2w_bench_u$u.err").read()[-1500:])
PY
done
