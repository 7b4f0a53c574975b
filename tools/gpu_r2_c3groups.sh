#!/bin/bash
for g in 1 2 3 4 5 6 8; do
  echo -n "groups $g: "; BOSS_LL_GROUPS=$g timeout 300 python tools/bench_configs.py --configs c3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['kernel'], round(j['ms_per_step'],4), round(j['frac_of_dgemm_peak'],4), end='  ')
print()"
done
echo -n "full job (S=4096 on one GPU), groups 4: "; timeout 300 python tools/bench_configs.py --configs c3 --full 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['kernel'], j['S_per_gpu'], round(j['ms_per_step'],4), round(j['frac_of_dgemm_peak'],4), end='  ')
print()"
