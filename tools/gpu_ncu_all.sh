#!/bin/bash
# ncu --set full captures of every kernel class on the two headline paths (one gpurun call, one GPU).
# Each ncu command follows a plain run of the same command that exited 0.  Outputs: gpurun_out/*.ncu-rep
mkdir -p gpurun_out
set -x
python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/plain_ll.log 2>&1 || exit 1
# loglik step = 4 stream groups x 48 launches (build_k, 16x potrf, 15x diagonal update, trsm + 14x panel, finish)
ncu --set full --clock-control none --import-source on -s 165 -c 9 -o gpurun_out/prof_ll_mid \
    python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_ll_mid.log 2>&1
echo "ncu ll mid exit $?"
ncu --set full --clock-control none --import-source on -k regex:'build_k|loglik_finish' -s 6 -c 2 -o gpurun_out/prof_ll_ends \
    python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_ll_ends.log 2>&1
echo "ncu ll ends exit $?"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/plain_sc.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'xcov|acq_kernel|argmax_final' -s 12 -c 3 -o gpurun_out/prof_score_small \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_sc_small.log 2>&1
echo "ncu score small exit $?"
ls -la gpurun_out/*.ncu-rep
