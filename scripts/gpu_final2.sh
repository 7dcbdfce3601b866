#!/bin/bash
# round-2 final single-GPU evidence (after the binning rework): GPU test suite, bench lines of every BASELINE config,
# CPU reference arm, smoke, ncu launch list + full captures, compute-sanitizer logs
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
[ -n "$SKIP_REFERENCE" ] || timeout 300 python bench.py --impl reference --steps 4 --warmup 1 --cpu-budget-s 45 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
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
# ncu: launch list of a bench run (eager launches: ncu lists kernels launched from the host), then full captures
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vcr --graph off > gpurun_out/r2_bench_under_ncu.log 2>&1
echo "launch list rc $?"
for k in ${NCU_KERNELS:-render_fwd_kernel render_bwd_transposed_kernel preprocess_fwd_kernel preprocess_bwd_kernel radix_onesweep_kernel emit_kernel}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -o gpurun_out/r2_full_$k -f python scripts/perf_probe.py --iters 1 > gpurun_out/r2_ncu_$k.log 2>&1
  echo "$k rc $?"
done
[ -n "$SKIP_GATHER4" ] || GSB_FWD_VARIANT=gather4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_kernel -s 12 -c 1 -o gpurun_out/r2_full_render_fwd_gather4 -f python scripts/perf_probe.py --iters 1 > gpurun_out/r2_ncu_render_fwd_gather4.log 2>&1
echo "gather4 capture rc $?"
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r2_sanitizer_$tool.log python scripts/sanitize_probe.py > gpurun_out/r2_sanitizer_$tool.out 2>&1
  echo "$tool rc $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2_sanitizer_$tool.log
done
ls -la gpurun_out/r2_full_*.ncu-rep | awk '{print $5, $9}'
