#!/bin/bash
# round-2 GPU pass 14: keys per block of the onesweep radix passes (length of the look-back chain)
mkdir -p gpurun_out
run_bench() {  # name, extra build flags
  GSB_NVCC_EXTRA="$2" python -m gaussianip_b200.build > /dev/null 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2m_bench_$1.json 2> gpurun_out/r2m_bench_$1.err
  echo "$1 rc $?"
}
run_bench ipt8 ""
run_bench ipt16 "-DGSB_RADIX_IPT=16"
run_bench ipt12 "-DGSB_RADIX_IPT=12"
GSB_NVCC_EXTRA="-DGSB_RADIX_IPT=16" python -m gaussianip_b200.build > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_knn.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "forward_exact or knn" > gpurun_out/r2m_tests.txt 2>&1
tail -3 gpurun_out/r2m_tests.txt
python -m gaussianip_b200.build > /dev/null 2>&1
python - <<'PY'
import json
for v in ("ipt8","ipt16","ipt12"):
    try:
        d=json.loads(open(f"gpurun_out/r2m_bench_{v}.json").read().strip().splitlines()[-1])
        st=d["roofline"]["stage_us_per_view"]
        print(v, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k: st[k] for k in ("depth_sort","scan_emit","tile_sort","ranges")})
    except Exception as e:
        print(v, "ERR", e)
PY
