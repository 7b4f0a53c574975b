#!/bin/bash
# full GPU test suite + bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_full.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2_pytest_full.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
echo "bench exit $?"; tail -3 gpurun_out/r2_bench_c.err; python tools/show_bench.py gpurun_out/r2_bench_c.json 2>/dev/null | head -40
python - <<'P'
import json
j=json.loads(open('gpurun_out/r2_bench_c.json').read().strip().splitlines()[-1])
print('launches/step', j['gpu_launches']/j['steps'], 'e2e', j['e2e']['value'], 'pageable', j['e2e']['pageable']['value'])
for c in j['configs']:
    if c['config']=='C4': print(json.dumps(c['on_device_multistart']))
P
