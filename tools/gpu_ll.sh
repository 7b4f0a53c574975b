#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
python bench.py --only loglik --steps 3 --warmup 3 > gpurun_out/ll.json 2>gpurun_out/ll.err; cat gpurun_out/ll.json; tail -3 gpurun_out/ll.err
python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/plain_ll.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 96 --csv --log-file gpurun_out/launches_ll.csv \
   python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_ll.log 2>&1
echo "ncu exit $?"
