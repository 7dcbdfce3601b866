#!/bin/bash
# round-2 GPU pass 6: transposed two-phase replay backward: parity, bench, variants
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graph.py tests/test_gpu_configs.py -m gpu -q --timeout 1200 -p no:cacheprovider -x > gpurun_out/r2f_tests.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r2f_tests.txt
tail -6 gpurun_out/r2f_tests.txt
for v in native replay_bwd; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr --variant $v > gpurun_out/r2f_bench_$v.json 2> gpurun_out/r2f_bench_$v.err
  echo "$v rc $?"
done
python scripts/perf_probe.py --iters 5 > gpurun_out/r2f_probe.txt 2>&1
tail -10 gpurun_out/r2f_probe.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd_transposed_kernel -s 12 -c 1 -o gpurun_out/r2f_render_bwd_transposed -f python scripts/perf_probe.py --iters 1 > gpurun_out/r2f_ncu_bwd.log 2>&1
python - <<'PY'
import json
for v in ("native","replay_bwd"):
    try:
        d=json.loads(open(f"gpurun_out/r2f_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), (d.get("roofline") or {}).get("stage_us_per_view"))
    except Exception as e:
        print(v, "ERR", e)
PY
