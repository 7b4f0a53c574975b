#!/bin/bash
# quick regression: GPU tests + the two loglik shapes
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -8
python bench.py --only loglik --steps 5 --warmup 3
python tools/bench_configs.py --configs c3 --steps 10 | cut -c1-330
