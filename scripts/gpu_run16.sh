#!/bin/bash
# round-2 GPU pass 16: staggered start of the views' forward pipelines
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
for k in 0 1 2; do
  GSB_FWD_STAGGER=$k timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2o_bench_s$k.json 2> gpurun_out/r2o_bench_s$k.err
  echo "s$k rc $?"
done
GSB_FWD_STAGGER=1 timeout 300 python bench.py --config vcr --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_vcr_s1.json 2> gpurun_out/r2o_bench_vcr_s1.err
GSB_FWD_STAGGER=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graph.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "batched or multistream or graph or fused_activations" > gpurun_out/r2o_tests.txt 2>&1
tail -3 gpurun_out/r2o_tests.txt
python - <<'PY'
import json
for v in ("s0","s1","s2","vcr_s1"):
    try:
        d=json.loads(open(f"gpurun_out/r2o_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
    except Exception as e:
        print(v, "ERR", e)
PY
