#!/bin/bash
# round-2 GPU pass 9: full GPU suite on the current tree, per-Gaussian backward grouping sweep, config lines
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r2i_tests.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r2i_tests.txt
tail -5 gpurun_out/r2i_tests.txt
for g in 8 1 2; do
  GSB_PBWD_GROUP=$g timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vcr > gpurun_out/r2i_bench_g$g.json 2> gpurun_out/r2i_bench_g$g.err
  echo "g$g rc $?"
done
timeout 300 python bench.py --config vcr --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_vcr.json 2> gpurun_out/r2i_bench_vcr.err
GSB_PBWD_GROUP=1 timeout 300 python bench.py --config vcr --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_vcr_g1.json 2> gpurun_out/r2i_bench_vcr_g1.err
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_c3.json 2> gpurun_out/r2i_bench_c3.err
timeout 300 python bench.py --config playback --steps 60 --warmup 6 --no-cpu-baseline > gpurun_out/r2i_bench_playback.json 2> gpurun_out/r2i_bench_playback.err
echo "playback rc $?"; tail -2 gpurun_out/r2i_bench_playback.err | cut -c1-300
python - <<'PY'
import json
for v in ("g8","g1","g2","vcr","vcr_g1","c3","playback"):
    try:
        d=json.loads(open(f"gpurun_out/r2i_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["value"],1), d["unit"], "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), (d.get("roofline") or {}).get("stage_us_per_view"))
    except Exception as e:
        print(v, "ERR", e)
PY
