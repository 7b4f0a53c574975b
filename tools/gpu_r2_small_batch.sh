#!/bin/bash
# small-batch kernels: latency table (wide / narrow / quarter / auto, bitwise comparison), the invariance tests, C4
mkdir -p gpurun_out
SB_SIZES=${SB_SIZES:-1,64,96,128,256,512,1024} timeout 900 python tools/bench_small_batch.py > gpurun_out/small_batch.jsonl 2> gpurun_out/small_batch.err
echo "small-batch exit $?"; tail -3 gpurun_out/small_batch.err
python - <<'P'
import json
for l in open('gpurun_out/small_batch.jsonl'):
    j=json.loads(l)
    print(j['n'], j['batch'], 'vg: wide', j['wide_vg_us'], 'narrow', j['narrow_vg_us'], 'quarter', j['quarter_vg_us'], 'auto', j['auto_vg_us'],
          ' v:', j['wide_v_us'], j['narrow_v_us'], j['quarter_v_us'], j['auto_v_us'], 'bits equal:', all(v for k, v in j.items() if 'bits' in k))
P
timeout 600 python -m pytest tests -m gpu -x -q -k "bitwise or split or multistart or tiny or repeatable" > gpurun_out/pytest_small.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_small.log
timeout 300 python tools/bench_configs.py --configs c4 2>/dev/null | head -1 | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('c4:', j['on_device_multistart'])"
