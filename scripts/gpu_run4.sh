#!/bin/bash
# round-2 GPU pass 4: ncu full captures of the main kernels + a no-reduction diagnostic timing
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
python scripts/perf_probe.py --iters 5 > gpurun_out/r2d_probe.txt 2>&1
for k in render_bwd_replay_kernel render_fwd_kernel radix_onesweep_kernel preprocess_fwd_kernel preprocess_bwd_kernel emit_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -o gpurun_out/r2d_$k -f python scripts/perf_probe.py --iters 1 > gpurun_out/r2d_ncu_$k.log 2>&1
  echo "$k rc $?"
done
GSB_NVCC_EXTRA="-DGSB_BWD_NO_RED" python -m gaussianip_b200.build > /dev/null 2>&1
python scripts/perf_probe.py --iters 5 > gpurun_out/r2d_probe_nored.txt 2>&1
python -m gaussianip_b200.build > /dev/null 2>&1
cat gpurun_out/r2d_probe.txt | tail -12; grep render_bwd gpurun_out/r2d_probe_nored.txt
ls -la gpurun_out/*.ncu-rep
