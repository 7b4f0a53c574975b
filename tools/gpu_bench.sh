#!/bin/bash
# bench + ncu evidence in one gpurun call.  Outputs in gpurun_out/.
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
tail -c 3000 gpurun_out/bench.json
tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_launch.log 2>&1
  echo "ncu launches exit $?"
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:score_trmm -s 8 -c 2 -o gpurun_out/prof_score \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
