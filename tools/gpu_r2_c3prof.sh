#!/bin/bash
mkdir -p gpurun_out
echo "groups 1"; BOSS_LL_GROUPS=1 timeout 300 python tools/bench_configs.py --configs c3 2>&1 | cut -c1-120
timeout 300 python tools/bench_configs.py --configs c3 > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3_r2.csv python tools/bench_configs.py --configs c3 > gpurun_out/ncu_c3_r2.log 2>&1
echo "ncu exit $?"; python tools/summarize_launches.py gpurun_out/launches_c3_r2.csv | head -20
