#!/bin/bash
mkdir -p gpurun_out
SB_SIZES=1,256 timeout 600 python tools/bench_small_batch.py > gpurun_out/small_batch4.jsonl 2> gpurun_out/small_batch4.err
echo "small-batch exit $?"; tail -3 gpurun_out/small_batch4.err; cut -c1-2000 gpurun_out/small_batch4.jsonl | sed 's/"wide_v_us/\n   &/; s/"quarter_v_us/\n   &/; s/"auto_v_us/\n   &/'
timeout 300 python tools/bench_configs.py --configs c4 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('c4:', j['on_device_multistart'])"
