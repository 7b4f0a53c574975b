#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py --configs c4 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('c4:', j['on_device_multistart'], j['ms_per_value_grad_iteration'])"
SB_SIZES=1,64,256,512 timeout 600 python tools/bench_small_batch.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['n'], j['batch'], 'wide', j['wide_vg_us'], 'narrow', j['narrow_vg_us'], 'quarter', j['quarter_vg_us'], 'auto', j['auto_vg_us'], 'v: ', j['wide_v_us'], j['narrow_v_us'], j['quarter_v_us'], j['auto_v_us'])
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4b.csv python tools/bench_configs.py --configs c4 --steps 1 > gpurun_out/ncu_c4b.log 2>&1
echo "ncu exit $?"
