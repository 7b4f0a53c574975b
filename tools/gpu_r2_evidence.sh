#!/bin/bash
# Round-2 evidence pass: full GPU test suite, bench line (+ reference arm), launch lists of the scoring and log-likelihood steps.
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_$tag.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench exit $?"; tail -3 gpurun_out/bench_$tag.err; python tools/show_bench.py gpurun_out/bench_$tag.json 2>/dev/null | head -60
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
echo "reference arm exit $?"; cut -c1-400 gpurun_out/bench_ref_$tag.json
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/plain_sc.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_$tag.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
timeout 300 python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/plain_ll.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_loglik_$tag.csv \
    python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_ll.log 2>&1
echo "ncu loglik launches exit $?"
ls -la gpurun_out/*_$tag*
