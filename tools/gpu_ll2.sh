#!/bin/bash
mkdir -p gpurun_out
python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/plain_ll.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 96 --csv --log-file gpurun_out/launches_ll.csv \
   python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_ll.log 2>&1
echo "ncu exit $?"
