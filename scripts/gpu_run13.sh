#!/bin/bash
# round-2 GPU pass 13: final ncu evidence: launch list of a bench run + full captures of every hot-path kernel
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vcr --graph off > gpurun_out/r2_bench_under_ncu.log 2>&1
echo "launch list rc $?"
for k in render_fwd_kernel render_bwd_transposed_kernel preprocess_fwd_kernel preprocess_bwd_kernel radix_onesweep_kernel emit_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -o gpurun_out/r2_full_$k -f python scripts/perf_probe.py --iters 1 > gpurun_out/r2_ncu_$k.log 2>&1
  echo "$k rc $?"
done
ls -la gpurun_out/r2_full_*.ncu-rep
