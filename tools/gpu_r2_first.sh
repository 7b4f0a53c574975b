#!/bin/bash
# Round-2 GPU call 1: does the refactored library load and pass, what does the bench say, memcheck on the subset.
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1
tail -2 gpurun_out/r2_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/r2_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
echo "bench exit $?"; tail -3 gpurun_out/r2_bench_a.err; python tools/show_bench.py gpurun_out/r2_bench_a.json 2>/dev/null | head -40
BOSS_UNFUSED_DIAG=1 timeout 300 python tools/bench_configs.py --configs c3 > gpurun_out/r2_c3_unfused.jsonl 2>&1
echo "c3 unfused:"; cut -c1-400 gpurun_out/r2_c3_unfused.jsonl
BOSS_UNFUSED_DIAG=1 timeout 300 python bench.py --only loglik --steps 3 --warmup 3 2>/dev/null | tail -1
timeout 300 python bench.py --only loglik --steps 3 --warmup 3 2>/dev/null | tail -1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitizer_subset.py > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck exit $?"; tail -8 gpurun_out/r2_memcheck.log
