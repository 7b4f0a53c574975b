#!/bin/bash
# A/B: new DMMA micro-tile diagonal-block kernel vs the previous scalar one (BOSS_POTRF_V1=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15
echo "--- new"; python bench.py --only loglik --steps 3 --warmup 3; python tools/bench_configs.py --configs c3 --steps 5 | cut -c1-330
echo "--- old"; BOSS_POTRF_V1=1 python bench.py --only loglik --steps 3 --warmup 3; BOSS_POTRF_V1=1 python tools/bench_configs.py --configs c3 --steps 5 | cut -c1-330
