#!/bin/bash
mkdir -p gpurun_out
for v in "" "BOSS_MS_NO_FAN=1"; do
  echo "== $v"; env $v timeout 300 python tools/diag/ms_multi_diag.py 2>&1 | tail -4
done
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/pytest_r2_n2.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_r2_n2.log
timeout 300 python tools/bench_configs.py --configs c4 2>/dev/null | head -1 | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('c4:', j['on_device_multistart'])"
timeout 600 python tools/bench_multi_inlib.py > gpurun_out/inlib_multi_n2.json 2>/dev/null; cut -c1-900 gpurun_out/inlib_multi_n2.json
timeout 600 python tools/parity_report.py > gpurun_out/parity_report.json 2> gpurun_out/parity_report.err; echo "parity exit $?"; head -20 gpurun_out/parity_report.json
