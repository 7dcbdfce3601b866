#!/bin/bash
# round-2 GPU pass 17: two-level radix look-back, fused scan+emit, tile ranges from the last sort pass
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 300 python scripts/sort_bench.py > gpurun_out/r2p_sort.txt 2>&1; echo "sort rc $?"
grep -c "correct=True" gpurun_out/r2p_sort.txt; grep "correct=False" gpurun_out/r2p_sort.txt | head
grep "n=  1048576\|n=  2097152" gpurun_out/r2p_sort.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r2p_tests.txt 2>&1
echo "pytest rc $?"; tail -4 gpurun_out/r2p_tests.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
echo "bench rc $?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2p_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "vcr", round(d.get("vcr",{}).get("value",0),1))
    print(d["roofline"]["stage_us_per_view"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r2p_bench.err").read()[-2000:])
PY
