#!/bin/bash
# round-2 GPU pass 7: transposed (two-phase) forward blend: whole parity suite on it, then A/B timings
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
GSB_FWD_VARIANT=transposed timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graph.py tests/test_gpu_configs.py tests/test_gpu_renderers.py tests/test_gpu_golden.py -m gpu -q --timeout 1200 -p no:cacheprovider -x > gpurun_out/r2g_tests.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r2g_tests.txt
tail -6 gpurun_out/r2g_tests.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2g_bench_perhit.json 2> gpurun_out/r2g_bench_perhit.err
GSB_FWD_VARIANT=transposed timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2g_bench_transposed.json 2> gpurun_out/r2g_bench_transposed.err
GSB_FWD_VARIANT=transposed python scripts/perf_probe.py --iters 5 > gpurun_out/r2g_probe.txt 2>&1
tail -10 gpurun_out/r2g_probe.txt
GSB_FWD_VARIANT=transposed timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_transposed_kernel -s 12 -c 1 -o gpurun_out/r2g_render_fwd_transposed -f python scripts/perf_probe.py --iters 1 > gpurun_out/r2g_ncu_fwd.log 2>&1
python - <<'PY'
import json
for v in ("perhit","transposed"):
    try:
        d=json.loads(open(f"gpurun_out/r2g_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), (d.get("roofline") or {}).get("stage_us_per_view"))
    except Exception as e:
        print(v, "ERR", e)
PY
