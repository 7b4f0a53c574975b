#!/bin/bash
mkdir -p gpurun_out
SB_SHAPES=1920x8,2048x8,2176x8,3968x4,4096x4,4224x4 SB_SIZES=1 timeout 600 python tools/bench_small_batch.py > gpurun_out/small_batch6.jsonl 2> gpurun_out/small_batch6.err
echo "exit $?"; tail -3 gpurun_out/small_batch6.err; python - <<'P'
import json
for l in open('gpurun_out/small_batch6.jsonl'):
    j=json.loads(l); print(j['n'], 'quarter v', j['quarter_v_us'], 'vg', j['quarter_vg_us'], 'narrow vg', j['narrow_vg_us'], j['auto_vg_kernel_us'])
P
