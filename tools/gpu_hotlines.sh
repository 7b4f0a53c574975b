#!/bin/bash
# Per-source-line stall samples (ncu --set full --import-source on) of the two dominant kernels; the source pages are
# exported on the box (the reports themselves are too large to bring back together).  Output: gpurun_out/hot_*.txt
mkdir -p gpurun_out
python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/plain_ll.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:chol_panel -s 190 -c 1 -o gpurun_out/prof_panel -f \
    python bench.py --only loglik --steps 1 --warmup 3 > gpurun_out/ncu_panel.log 2>&1
echo "ncu panel exit $?"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/plain_sc.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:score_trmm -s 8 -c 1 -o gpurun_out/prof_trmm -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_trmm.log 2>&1
echo "ncu trmm exit $?"
for r in prof_panel prof_trmm; do
  ncu -i gpurun_out/$r.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$r.src.csv 2>/dev/null
  python tools/ncu_source_hot.py gpurun_out/$r.src.csv 30 > gpurun_out/hot_$r.txt 2>&1
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep gpurun_out/$r.src.csv
done
head -40 gpurun_out/hot_prof_panel.txt
