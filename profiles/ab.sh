#!/bin/bash
# A/B experimental builds on the GPU box: ab.sh <lib.so> [bench args]; prints time + instruction counts
lib=$1; shift
RTB200_LIB=$PWD/$lib bash profiles/quick_bench.sh "$*"
RTB200_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:rt_batch_kernel -s 3 -c 1 python bench.py --steps 2 --warmup 1 --no-cpu $* 2>&1 | grep -E "duration|inst_executed|fp64|issue_active" | awk '{printf "%s=%s ", $1, $NF} END {print ""}'
