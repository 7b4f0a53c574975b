#!/bin/bash
mkdir -p gpurun_out
for g in 2 4 6 8 12 16; do
  echo "groups $g"; BOSS_LL_GROUPS=$g timeout 300 python tools/bench_configs.py --configs c3 2>&1 | cut -c1-200 | sed 's/"n": 512.*"ms_per_step"/ms_per_step/' 
done
for g in 4 8; do echo "headline groups $g"; BOSS_LL_GROUPS=$g timeout 300 python bench.py --only loglik --steps 3 --warmup 3 2>/dev/null | tail -1; done
