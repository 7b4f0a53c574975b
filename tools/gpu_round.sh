#!/bin/bash
# One gpurun call: FP64 ceilings, parity tests, smoke.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
if [ "$1" != "nomicro" ]; then
  timeout 120 ./tools/microbench.bin > gpurun_out/microbench.json 2> gpurun_out/microbench.err
  timeout 300 python tools/dgemm_peak.py > gpurun_out/dgemm_peak.json 2> gpurun_out/dgemm_peak.err
fi
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
cat gpurun_out/microbench.json gpurun_out/dgemm_peak.json 2>/dev/null
