#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "bitwise or split or multistart or redzone or tiny or repeatable or host_mirror" > gpurun_out/pytest_small3.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_small3.log
BOSS_MS_TRACE=1 timeout 300 python tools/bench_configs.py --configs c4 > gpurun_out/c4_fan.jsonl 2> gpurun_out/c4_fan.err
echo "c4 exit $?"; grep -c "round" gpurun_out/c4_fan.err; grep "evaluated" gpurun_out/c4_fan.err | tail -1; cut -c1-900 gpurun_out/c4_fan.jsonl
grep round gpurun_out/c4_fan.err | awk '{print $6"x"$8}' | tr '\n' ' ' | cut -c1-1500
BOSS_MS_NO_FAN=1 timeout 300 python tools/bench_configs.py --configs c4 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('no fan:', j['on_device_multistart'])"
