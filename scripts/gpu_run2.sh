#!/bin/bash
# round-2 GPU pass 2: hit-record forward + replay backward: parity, then timings of the variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graph.py tests/test_gpu_renderers.py tests/test_gpu_adam.py -m gpu -q -rA --timeout 900 -p no:cacheprovider -x > gpurun_out/r2b_tests.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r2b_tests.txt
tail -5 gpurun_out/r2b_tests.txt
for v in native packed_bwd rescan_bwd; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr --variant $v > gpurun_out/r2b_bench_$v.json 2> gpurun_out/r2b_bench_$v.err
  echo "$v rc $?"; tail -2 gpurun_out/r2b_bench_$v.err
done
timeout 300 python bench.py --config playback --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_playback.json 2> gpurun_out/r2b_bench_playback.err
echo "playback rc $?"; tail -3 gpurun_out/r2b_bench_playback.err
python - <<'PY'
import json
for v in ("native","packed_bwd","rescan_bwd","playback"):
    try:
        d=json.loads(open(f"gpurun_out/r2b_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), (d.get("roofline") or {}).get("stage_us_per_view"))
    except Exception as e:
        print(v, "ERR", e)
PY
