#!/bin/bash
# round-2 final single-GPU evidence: GPU test suite, the bench lines of every BASELINE config, the CPU reference arm
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r2_gpu_tests.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r2_gpu_tests.txt
tail -4 gpurun_out/r2_gpu_tests.txt
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc $?"
timeout 300 python bench.py --config vcr --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_vcr.json 2> gpurun_out/r2_bench_vcr.err
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err
timeout 300 python bench.py --config playback --steps 136 --warmup 8 --no-cpu-baseline > gpurun_out/r2_bench_playback.json 2> gpurun_out/r2_bench_playback.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 --cpu-budget-s 45 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -1 gpurun_out/r2_smoke.txt
python - <<'PY'
import json
for v in ("","_vcr","_c3","_playback","_reference"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench{v}.json").read().strip().splitlines()[-1])
        print(v or "default", round(d["value"],3), d["unit"], "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"), (d.get("vcr") or {}).get("value"))
    except Exception as e:
        print(v, "ERR", e)
PY
