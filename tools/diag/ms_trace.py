import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import boss_b200
from boss_b200 import _lib
from tests.util_problems import make_problem
_lib.init(0)
n, d, M = 4096, 4, 1024
X, Y, ls, amp, ns = make_problem(n, d, seed=1004)
theta = np.array([0.8, -0.3]); mean_X = theta[0] * X[0] + theta[1]; y = Y[0] + mean_X
gp = _lib.gp_fit(X, y - mean_X, ls[0], amp[0], ns[0], 2)
best = float(np.max(y))
rng = np.random.default_rng(4004)
starts = (rng.permuted(np.tile(np.arange(M), (d, 1)), axis=1) + rng.random((d, M))) / M
aff = np.zeros((1, d + 1)); aff[0, 0] = theta[1]; aff[0, 1] = theta[0]
lb, ub = np.zeros(d), np.ones(d)
for rep in range(2):
    t0 = time.perf_counter()
    r = _lib.ei_maximize_multistart([gp], 1, 1, starts, [1.0], best, None, lb, ub, iters=50, prior_mean_affine=aff)
    print("wall", time.perf_counter() - t0, "evals", r[5], "best", r[3], flush=True)
# cost of small value+grad evaluations
import torch
for m in (1, 32, 100, 300, 1024):
    Xs = starts[:, :m]
    _lib.ei_value_grad([gp], 1, 1, Xs, [1.0], best, None, lb=lb, ub=ub)
    t0 = time.perf_counter()
    for _ in range(20):
        _lib.ei_value_grad([gp], 1, 1, Xs, [1.0], best, None, lb=lb, ub=ub)
    print("value_grad host call m=%d: %.3f ms" % (m, (time.perf_counter() - t0) / 20 * 1e3), flush=True)
