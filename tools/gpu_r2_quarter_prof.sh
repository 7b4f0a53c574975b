#!/bin/bash
# ncu --set full + per-line stalls of the tiny-batch kernels (n = 4096 instance: the second problem of bench_small_batch)
mkdir -p gpurun_out
SB_SIZES=8 timeout 300 python tools/bench_small_batch.py > gpurun_out/plain_small.log 2>&1 || exit 1
SB_SIZES=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_quarter -s 200 -c 2 -o gpurun_out/prof_quarter -f \
    python tools/bench_small_batch.py > gpurun_out/ncu_quarter.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_quarter.log
ncu -i gpurun_out/prof_quarter.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_quarter.src.csv 2>/dev/null
python tools/ncu_source_hot.py gpurun_out/prof_quarter.src.csv 40 > gpurun_out/hot_prof_quarter.txt 2>&1
ncu -i gpurun_out/prof_quarter.ncu-rep --page raw --csv > gpurun_out/prof_quarter.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_quarter.ncu-rep --page details > gpurun_out/prof_quarter.details.txt 2>/dev/null
rm -f gpurun_out/prof_quarter.ncu-rep gpurun_out/prof_quarter.src.csv
python tools/ncu_summary.py gpurun_out/prof_quarter.raw.csv
head -50 gpurun_out/hot_prof_quarter.txt
