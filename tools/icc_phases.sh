#!/bin/bash
# instruction-cache requests / time of one update-kernel launch: fused, physics only, generator only; jump on / off
for jump in 1 0; do for skip in 0 1 2; do
  echo "== JUMP=$jump SKIP=$skip"
  QCDGPU_RLX_JUMP=$jump QCDGPU_DEBUG_SKIP=$skip ncu --metrics sm__icc_requests.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg \
    --clock-control none -k regex:update_kernel_rlx -s 16 -c 1 python bench.py --steps 1 --warmup 2 --no-cpu 2>&1 | grep -E "icc|duration|inst_executed|no_instruction|issue_active|cycles_elapsed" | awk '{printf "%s=%s  ", $1, $NF} END{print ""}'
done; done
