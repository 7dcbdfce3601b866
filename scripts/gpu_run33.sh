#!/bin/bash
# final sanity on the final tree: GPU suite, default bench line, smoke
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2_gpu_tests.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r2_gpu_tests.txt; tail -3 gpurun_out/r2_gpu_tests.txt
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc $?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -1 gpurun_out/r2_smoke.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench.json").read().strip().splitlines()[-1])
print("default", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["roofline"]["kernel"], round(d["roofline"]["frac"],4), "vcr", round(d["vcr"]["value"],1), "cpu", d["cpu_baseline"]["value"], d["clocks"])
print(d["roofline"]["stage_us_per_view"])
PY
