#!/bin/bash
# whole-step counters: ncu profiles each replayed CUDA graph (one 4-view step) as ONE workload
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
timeout 900 ncu --graph-profiling graph --clock-control none \
  --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
  --csv --log-file gpurun_out/r2s_graph_step.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-vcr > gpurun_out/r2s_bench_under_ncu.log 2>&1
echo "ncu rc $?"
grep -c . gpurun_out/r2s_graph_step.csv
grep -i "graph" gpurun_out/r2s_graph_step.csv | head -40 | cut -c1-300
tail -3 gpurun_out/r2s_bench_under_ncu.log | cut -c1-300
