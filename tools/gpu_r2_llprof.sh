#!/bin/bash
mkdir -p gpurun_out
python bench.py --only loglik --steps 1 --warmup 3 > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 110 --csv --log-file gpurun_out/launches_ll_fused_r2.csv python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_ll_r2.log 2>&1
echo "ncu fused exit $?"; python tools/summarize_launches.py gpurun_out/launches_ll_fused_r2.csv | head
BOSS_UNFUSED_DIAG=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 580 -c 160 --csv --log-file gpurun_out/launches_ll_unfused_r2.csv python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_ll_r2b.log 2>&1
echo "ncu unfused exit $?"; python tools/summarize_launches.py gpurun_out/launches_ll_unfused_r2.csv | head
