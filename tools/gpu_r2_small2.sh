#!/bin/bash
mkdir -p gpurun_out
SB_SIZES=1,32,256 timeout 600 python tools/bench_small_batch.py > gpurun_out/small_batch2.jsonl 2> gpurun_out/small_batch2.err
echo "small-batch exit $?"; tail -3 gpurun_out/small_batch2.err; cut -c1-2000 gpurun_out/small_batch2.jsonl | sed 's/"wide.*"auto_v_us"/"auto_v_us"/'
SB_SIZES=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_small.csv python tools/bench_small_batch.py > gpurun_out/ncu_small.log 2>&1
echo "ncu exit $?"
