#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "loglik or c3 or small_matrix or fit_batch or multistart or repeatable" > gpurun_out/r2_pytest3.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/r2_pytest3.log
timeout 300 python tools/bench_configs.py --configs c3,c4 > gpurun_out/r2_c3_permatrix.jsonl 2>&1
echo "c3 per-matrix:"; cut -c1-330 gpurun_out/r2_c3_permatrix.jsonl
BOSS_NO_PER_MATRIX=1 timeout 300 python tools/bench_configs.py --configs c3 > gpurun_out/r2_c3_fuseddiag.jsonl 2>&1
echo "c3 fused-diag multi-kernel:"; cut -c1-330 gpurun_out/r2_c3_fuseddiag.jsonl
