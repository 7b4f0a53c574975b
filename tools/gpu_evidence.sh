#!/bin/bash
# One gpurun call: the bench line plus every ncu artefact profiles/ cites, tagged $1 (e.g. v4).
# Each ncu command follows a plain run of the same command that exited 0; nothing printed under ncu is a bench value.
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
ncu --set full --clock-control none --import-source on -k regex:score_trmm -s 8 -c 2 -o gpurun_out/prof_score_$tag -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_full.log 2>&1
echo "ncu score_trmm exit $?"
ncu --set full --clock-control none --import-source on -k regex:'xcov|acq_kernel|argmax_final' -s 12 -c 3 -o gpurun_out/prof_score_small_$tag -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_sc_small.log 2>&1
echo "ncu score small exit $?"
python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/plain_ll.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 576 -c 192 --csv --log-file gpurun_out/launches_loglik_$tag.csv \
    python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_ll.log 2>&1
echo "ncu loglik launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:'chol_panel|chol_update|potrf_tile' -s 600 -c 9 -o gpurun_out/prof_ll_mid_$tag -f \
    python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_ll_mid.log 2>&1
echo "ncu ll mid exit $?"
ncu --set full --clock-control none --import-source on -k regex:'build_k|loglik_finish|chol_trsm' -s 20 -c 3 -o gpurun_out/prof_ll_ends_$tag -f \
    python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_ll_ends.log 2>&1
echo "ncu ll ends exit $?"
# gpurun copies back at most 64 MiB: export the pages read afterwards, keep only the score_trmm report itself
for r in prof_score_$tag prof_score_small_$tag prof_ll_mid_$tag prof_ll_ends_$tag; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  ncu -i gpurun_out/$r.ncu-rep --page source --csv > gpurun_out/$r.source.csv 2>/dev/null
done
rm -f gpurun_out/prof_score_small_$tag.ncu-rep gpurun_out/prof_ll_mid_$tag.ncu-rep gpurun_out/prof_ll_ends_$tag.ncu-rep
ls -la gpurun_out/*_$tag*
