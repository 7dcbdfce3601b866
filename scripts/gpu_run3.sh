#!/bin/bash
# round-2 GPU pass 3: chunk-level hit recording + RB-batched replay backward: parity, then RB sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r2c_tests.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r2c_tests.txt
tail -4 gpurun_out/r2c_tests.txt
for rb in 2 1 4; do
  GSB_NVCC_EXTRA="-DGSB_BWD_RB=$rb" python -m gaussianip_b200.build > /dev/null 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2c_bench_rb$rb.json 2> gpurun_out/r2c_bench_rb$rb.err
  echo "rb$rb rc $?"
done
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr --variant rescan_bwd > gpurun_out/r2c_bench_rescan.json 2> gpurun_out/r2c_bench_rescan.err
python - <<'PY'
import json
for v in ("rb2","rb1","rb4","rescan"):
    try:
        d=json.loads(open(f"gpurun_out/r2c_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), (d.get("roofline") or {}).get("stage_us_per_view"))
    except Exception as e:
        print(v, "ERR", e)
PY
