#!/bin/bash
# round-2 GPU pass 5: half-warp hit streams (forward + replay backward), hoisted per-Gaussian loads: parity, bench, sweeps
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 1500 -p no:cacheprovider > gpurun_out/r2e_tests.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r2e_tests.txt
tail -6 gpurun_out/r2e_tests.txt
run_bench() {  # name, extra build flags
  GSB_NVCC_EXTRA="$2" python -m gaussianip_b200.build > /dev/null 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2e_bench_$1.json 2> gpurun_out/r2e_bench_$1.err
  echo "$1 rc $?"
}
run_bench base ""
run_bench pbwd3 "-DGSB_PBWD_MINB=3"
run_bench fwdhb1 "-DGSB_FWD_HB=1"
python -m gaussianip_b200.build > /dev/null 2>&1
python scripts/perf_probe.py --iters 5 > gpurun_out/r2e_probe.txt 2>&1
tail -10 gpurun_out/r2e_probe.txt
python - <<'PY'
import json
for v in ("base","pbwd3","fwdhb1"):
    try:
        d=json.loads(open(f"gpurun_out/r2e_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), (d.get("roofline") or {}).get("stage_us_per_view"))
    except Exception as e:
        print(v, "ERR", e)
PY
