#!/bin/bash
mkdir -p gpurun_out
python -m gaussianip_b200.build > /dev/null 2>&1
GSB_FWD_VARIANT=precull timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_precull -s 12 -c 1 -o gpurun_out/r2_full_render_fwd_precull -f python scripts/perf_probe.py --iters 1 > gpurun_out/r2_ncu_render_fwd_precull.log 2>&1
echo "rc $?"
ncu -i gpurun_out/r2_full_render_fwd_precull.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
for k in ('gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','sm__cycles_active.avg','sm__cycles_elapsed.avg','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__thread_inst_executed_per_inst_executed.ratio'):
    print(k, v[h.index(k)])
"
