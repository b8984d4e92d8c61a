#!/bin/bash
# Issue-side metrics of one kernel in the one-frame-per-call path: tools/ncu_single.sh <out.csv> <kernel-regex> [skip] [lib]
OUT=$1; K=$2; SKIP=${3:-30}; LIB=$4
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_selected_per_issue_active.ratio,smsp__warps_active.avg.per_cycle_active
GOF_B200_LIB=$LIB timeout 300 ncu --metrics $M --clock-control none -k regex:$K -s $SKIP -c 1 --csv --log-file $OUT python tools/single_frame.py 256 3 > /dev/null 2>&1
python - $OUT <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1],errors="replace")) if len(r)>10]
h=rows[0]; ni,vi=h.index("Metric Name"),h.index("Metric Value")
for r in rows[1:]:
    print(f"  {r[ni]:90s} {r[vi]}")
PY
