#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "score or argmax or ei_ or predict or c5" > gpurun_out/r2_pytest_q.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_q.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
echo "bench exit $?"; python tools/show_bench.py gpurun_out/r2_bench_d.json 2>/dev/null | head -3
BOSS_UNFUSED_SCORE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2_bench_d2.json 2> gpurun_out/r2_bench_d2.err
echo "unfused:"; python tools/show_bench.py gpurun_out/r2_bench_d2.json 2>/dev/null | head -1
