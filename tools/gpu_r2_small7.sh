#!/bin/bash
mkdir -p gpurun_out
SB_SIZES=1,64,256,512 timeout 600 python tools/bench_small_batch.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['n'], j['batch'], 'vg: wide', j['wide_vg_us'], 'narrow', j['narrow_vg_us'], 'quarter', j['quarter_vg_us'], 'auto', j['auto_vg_us'], ' v:', j['wide_v_us'], j['narrow_v_us'], j['quarter_v_us'], j['auto_v_us'], all(v for k,v in j.items() if 'bits' in k))
"
timeout 300 python tools/bench_configs.py --configs c4 2>/dev/null | head -1 | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print(j['on_device_multistart'])"
