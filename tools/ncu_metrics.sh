#!/bin/bash
# Key issue-side metrics of one kernel on the headline workload: tools/ncu_metrics.sh <out.csv> <kernel-regex> [lib]
OUT=$1; K=$2; LIB=$3; EXTRA="${@:4}"
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,sm__cycles_active.avg,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
GOF_B200_LIB=$LIB timeout 600 ncu --metrics $M --clock-control none -k regex:$K -s ${SKIP:-8} -c 1 --csv --log-file $OUT python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline $EXTRA > /dev/null 2>&1
python - $OUT <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1],errors="replace")) if len(r)>10]
h=rows[0]; ni,vi=h.index("Metric Name"),h.index("Metric Value")
for r in rows[1:]:
    print(f"  {r[ni]:90s} {r[vi]}")
PY
