#!/bin/bash
# tools/icc.sh LIB...: instruction-cache requests, instructions and no-instruction stalls of two update-kernel launches per library
for L in "$@"; do
  echo "== $L"
  QCDGPU_LIB=$L ncu --metrics sm__icc_requests.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gcc__cache_requests_type_instruction.sum \
    --clock-control none -k regex:update_kernel_rlx -s 16 -c 2 python bench.py --steps 1 --warmup 2 --no-cpu 2>&1 | grep -E "update_kernel|icc|duration|inst_executed|no_instruction|issue_active|gcc" | cut -c1-160
done
