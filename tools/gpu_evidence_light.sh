#!/bin/bash
# Lighter variant of gpu_evidence.sh for changes that only touch the log-likelihood path: bench line, reference arm,
# both launch lists and the ncu --set full capture of the factorisation kernels, tagged $1.
tag=${1:-vX}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench exit $?"; tail -3 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
echo "reference arm exit $?"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/plain_sc.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 400 --csv --log-file gpurun_out/launches_bench_$tag.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/plain_ll.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 576 -c 192 --csv --log-file gpurun_out/launches_loglik_$tag.csv \
    python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_ll.log 2>&1
echo "ncu loglik launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:'chol_panel|chol_update|potrf_tile' -s 600 -c 9 -o gpurun_out/prof_ll_mid_$tag -f \
    python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_ll_mid.log 2>&1
echo "ncu ll mid exit $?"
ncu -i gpurun_out/prof_ll_mid_$tag.ncu-rep --page raw --csv > gpurun_out/prof_ll_mid_$tag.raw.csv 2>/dev/null
rm -f gpurun_out/prof_ll_mid_$tag.ncu-rep
ls -la gpurun_out/*_$tag*
