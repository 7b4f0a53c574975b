#!/bin/bash
# C3 / C4 diagnosis: multi-start round trace, launch lists of one C3 step and of the C4 on-device multi-start
mkdir -p gpurun_out
BOSS_MS_TRACE=1 timeout 300 python tools/bench_configs.py --configs c4 > gpurun_out/c4_trace.jsonl 2> gpurun_out/c4_trace.err
echo "c4 exit $?"; grep -c "round" gpurun_out/c4_trace.err; cut -c1-600 gpurun_out/c4_trace.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv \
    python tools/bench_configs.py --configs c4 --steps 1 > gpurun_out/ncu_c4.log 2>&1
echo "ncu c4 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv \
    python tools/bench_configs.py --configs c3 --steps 2 > gpurun_out/ncu_c3.log 2>&1
echo "ncu c3 exit $?"
