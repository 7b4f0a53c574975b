#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "multi_gpu or shutdown" > gpurun_out/r2_pytest_multi.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/r2_pytest_multi.log
timeout 600 python tools/bench_multi_inlib.py > gpurun_out/r2_inlib_multi.json 2> gpurun_out/r2_inlib_multi.err
echo "inlib exit $?"; tail -3 gpurun_out/r2_inlib_multi.err; cat gpurun_out/r2_inlib_multi.json
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "torchrun bench exit $?"; tail -2 gpurun_out/r2_bench_n$N.err; python tools/show_bench.py gpurun_out/r2_bench_n$N.json | head -12
