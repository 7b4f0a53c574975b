#!/bin/bash
# N-GPU evidence: single-process in-library path (boss_init_multi) and the torchrun bench line (with C3/C4/C5 per rank)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "multi_gpu" > gpurun_out/pytest_multi_n$N.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_multi_n$N.log
timeout 600 python tools/bench_multi_inlib.py > gpurun_out/inlib_multi_n$N.json 2> gpurun_out/inlib_multi_n$N.err
echo "inlib exit $?"; tail -3 gpurun_out/inlib_multi_n$N.err; cat gpurun_out/inlib_multi_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "torchrun bench exit $?"; tail -2 gpurun_out/bench_n$N.err; python tools/show_bench.py gpurun_out/bench_n$N.json | head -12
