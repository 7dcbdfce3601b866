#!/bin/bash
# round-2 GPU pass 10: lanes-per-Gaussian split of the fused per-Gaussian backward: parity + sweep; sanitizers
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graph.py tests/test_gpu_renderers.py -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r2j_tests.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r2j_tests.txt
tail -4 gpurun_out/r2j_tests.txt
run_bench() {  # name, extra build flags
  GSB_NVCC_EXTRA="$2" python -m gaussianip_b200.build > /dev/null 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2j_bench_$1.json 2> gpurun_out/r2j_bench_$1.err
  echo "$1 rc $?"
}
run_bench lpg4 ""
run_bench lpg2 "-DGSB_PBWD_LPG=2"
run_bench lpg1 "-DGSB_PBWD_LPG=1"
run_bench lpg4_b3 "-DGSB_PBWD_MINB=3"
python -m gaussianip_b200.build > /dev/null 2>&1
python - <<'PY'
import json
for v in ("lpg4","lpg2","lpg1","lpg4_b3"):
    try:
        d=json.loads(open(f"gpurun_out/r2j_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
    except Exception as e:
        print(v, "ERR", e)
PY
bash scripts/gpu_run9b.sh
