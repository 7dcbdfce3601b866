#!/bin/bash
# round-2 GPU pass 32: forward blend without any divergent region (cull test and record store branch-free too)
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r2y_tests_all2.txt 2>&1
echo "pytest rc $?"; tail -2 gpurun_out/r2y_tests_all2.txt
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2y_bench_nb3_$i.json 2> gpurun_out/r2y_bench_nb3_$i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2y_bench_nb3_$i.json").read().strip().splitlines()[-1])
s=d["roofline"]["stage_us_per_view"]
print("run $i value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "fwd", s["render_fwd"], "bwd", s["render_bwd"])
PY
done
