#!/bin/bash
# Round-2 ncu --set full captures (one GPU; each after a plain run of the same command exited 0): the fused scoring
# kernel + xcov, the log-likelihood factorisation kernels at n = 2048, the C3 (n = 512) kernels.  Raw CSV pages are exported
# on the box and summarised by tools/ncu_summary.py; the reports themselves stay on the box.
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/prof_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_$name.src.csv 2>/dev/null
  python tools/ncu_source_hot.py gpurun_out/prof_$name.src.csv 25 > gpurun_out/hot_$name.txt 2>&1
  rm -f gpurun_out/prof_$name.ncu-rep gpurun_out/prof_$name.src.csv
  python tools/ncu_summary.py gpurun_out/prof_$name.raw.csv
}
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/plain_sc.log 2>&1 || exit 1
cap trmm 'score_trmm' 8 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs
cap xcov 'xcov' 8 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs
timeout 300 python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/plain_ll.log 2>&1 || exit 1
cap ll 'chol_panel|chol_update_kernel|potrf_tile|build_k' 150 9 python bench.py --only loglik --steps 1 --warmup 3
timeout 300 python tools/bench_configs.py --configs c3 --steps 1 > gpurun_out/plain_c3.log 2>&1 || exit 1
cap c3 'potrf_fused|chol_panel|chol_trsm|build_k' 40 8 python tools/bench_configs.py --configs c3 --steps 1
ls -la gpurun_out/prof_*.raw.csv gpurun_out/hot_*.txt
