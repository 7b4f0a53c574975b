#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "multistart or score or predict or value_grad or repeatable or mirror or bo_loop" > gpurun_out/r2_pytest2.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2_pytest2.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
echo "bench exit $?"; tail -3 gpurun_out/r2_bench_b.err; python tools/show_bench.py gpurun_out/r2_bench_b.json 2>/dev/null | head -40
python - <<'P'
import json
j=json.loads(open('gpurun_out/r2_bench_b.json').read().strip().splitlines()[-1])
for c in j['configs']:
    if c['config']=='C4': print(json.dumps(c['on_device_multistart']))
P
