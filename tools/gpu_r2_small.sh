#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_small_batch.py > gpurun_out/small_batch.jsonl 2> gpurun_out/small_batch.err
echo "small-batch exit $?"; tail -3 gpurun_out/small_batch.err; cat gpurun_out/small_batch.jsonl
timeout 600 python -m pytest tests -m gpu -x -q -k "bitwise or split or multistart or redzone or small_batch" > gpurun_out/pytest_small.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_small.log
timeout 300 python tools/bench_configs.py --configs c4 > gpurun_out/c4_quarter.jsonl 2>&1
echo "c4 exit $?"; cut -c1-900 gpurun_out/c4_quarter.jsonl
