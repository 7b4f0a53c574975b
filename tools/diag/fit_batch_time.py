"""BI-style posterior construction: S fits one by one vs one boss_gp_fit_batch call."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import boss_b200
from boss_b200 import _lib
from tests.util_problems import make_problem, make_hyper_samples
_lib.init(0)
for n, d, S in ((512, 6, 64), (2048, 8, 40)):
    X, Y, _, _, _ = make_problem(n, d, seed=1)
    L, A, N = make_hyper_samples(S, d, seed=2)
    for rep in range(2):
        t0 = time.perf_counter(); gs = [_lib.gp_fit(X, Y[0], L[s], A[s], N[s], 2) for s in range(S)]; t1 = time.perf_counter() - t0
        [g.free() for g in gs if g]
        t0 = time.perf_counter(); gb = _lib.gp_fit_batch(X, Y[0], L, A, N, 2); t2 = time.perf_counter() - t0
        [g.free() for g in gb if g]
    print(f"n={n} d={d} S={S}: one by one {t1*1e3:.1f} ms, batched {t2*1e3:.1f} ms ({t1/t2:.1f}x)")
