#!/bin/bash
mkdir -p gpurun_out
BOSS_MS_TRACE=1 timeout 300 python tools/bench_configs.py --configs c4 > gpurun_out/c4_trace2.jsonl 2> gpurun_out/c4_trace2.err
grep round gpurun_out/c4_trace2.err | tail -100 | awk '{print $6"x"$8":"$10}' | tr '\n' ' '
