#!/bin/bash
# instruction-cache behaviour of the correction kernel versus residency (metrics-only ncu passes)
mkdir -p gpurun_out
: > gpurun_out/icc.log
M=sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,sm__warps_active.avg.per_cycle_active
for b in 1 2 4; do
  echo "== blocks/SM $b" >> gpurun_out/icc.log
  TALC_BLOCKS_PER_SM=$b ncu --metrics $M --clock-control none -k regex:correct_kernel -c 1 python tools/profile_case.py ${1:-8000} 1 2>&1 | grep -E "icc|inst_executed|duration|issue|warps_active" >> gpurun_out/icc.log
done
cat gpurun_out/icc.log
