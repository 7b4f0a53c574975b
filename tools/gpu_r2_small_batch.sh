#!/bin/bash
mkdir -p gpurun_out
SB_SIZES=1,64,256,1024 timeout 600 python tools/bench_small_batch.py > gpurun_out/small_batch5.jsonl 2> gpurun_out/small_batch5.err
echo "small-batch exit $?"; tail -3 gpurun_out/small_batch5.err; cut -c1-2000 gpurun_out/small_batch5.jsonl | sed 's/"wide_v_us/\n   &/; s/"quarter_v_us/\n   &/; s/"auto_v_us/\n   &/'
timeout 600 python -m pytest tests -m gpu -x -q -k "bitwise or split or multistart or tiny or repeatable" > gpurun_out/pytest_small5.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_small5.log
timeout 300 python tools/bench_configs.py --configs c4 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('c4:', j['on_device_multistart'])"
