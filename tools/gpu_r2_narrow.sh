#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "score or argmax or ei_ or predict or multistart or split or value_grad or mirror or bo_loop or edge or repeatable or cov" > gpurun_out/r2_pytest_n.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/r2_pytest_n.log
timeout 300 python tools/bench_configs.py --configs c4 > gpurun_out/r2_c4_narrow.jsonl 2>&1; cut -c1-900 gpurun_out/r2_c4_narrow.jsonl
BOSS_NO_NARROW=1 timeout 300 python tools/bench_configs.py --configs c4 > gpurun_out/r2_c4_wide.jsonl 2>&1; echo wide; cut -c1-900 gpurun_out/r2_c4_wide.jsonl
