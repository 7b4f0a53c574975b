"""Tiny cases through the per-matrix persistent Cholesky (chol_matrix_kernel) with progress prints."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import boss_b200
from boss_b200 import _lib
from oracle import boss_oracle as O
from tests.util_problems import make_hyper_samples, make_problem, relerr
_lib.init(0)
for n, d, kid, S in [(100, 3, 2, 1), (100, 3, 2, 3), (130, 3, 1, 1), (130, 3, 1, 17), (300, 4, 0, 5), (512, 6, 2, 24), (512, 6, 2, 700)]:
    X, Y, _, _, _ = make_problem(n, d, seed=300 + n)
    L, A, N = make_hyper_samples(S, d, seed=301 + n)
    t0 = time.time()
    out = _lib.loglik_batch(X, Y[0], L, A, N, kid)
    k = min(S, 24)
    ref = O.gp_loglik_batch(X, Y[0], L[:k], A[:k], N[:k], kid)
    print(n, d, kid, S, "relerr %.2e" % relerr(out[:k], ref), "%.3f s" % (time.time() - t0), flush=True)
