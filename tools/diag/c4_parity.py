"""Diagnostic: where does the C4 (n=4096, d=4) EI discrepancy vs the oracle come from?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import boss_b200
from boss_b200 import _lib
from oracle import boss_oracle as O
from tests.util_problems import make_problem
_lib.init(0)
n, d, M = 4096, 4, 2048
X, Y, ls, amp, ns = make_problem(n, d, seed=1004)
gp = _lib.gp_fit(X, Y[0], ls[0], amp[0], ns[0], 2)
post = O.posterior_fit(X, Y[0], ls[0], amp[0], ns[0], 2)
Xs = np.random.default_rng(4004).random((d, M))
mu, var, st = _lib.gp_predict(gp, Xs)
mu_r, var_r, _ = O.mean_and_var(post, Xs)
print("mu relerr max", np.max(np.abs(mu - mu_r) / np.abs(mu_r)), " var relerr max", np.max(np.abs(var - var_r) / var_r))
best = float(np.max(Y[0]))
acq, _, _ = _lib.ei_score([gp], 1, 1, Xs, [1.0], best, None)
ref, _, _ = O.ei_acquisition([[post]], Xs, [1.0], best, None)
z = (mu_r - best) / np.sqrt(var_r)
m = ref > 1e-200
rel = np.abs(acq[m] - ref[m]) / ref[m]
zz = z[m]
order = np.argsort(-rel)[:10]
for i in order:
    print(f"z={zz[i]:8.3f} EI={ref[m][i]:.3e} rel={rel[i]:.2e}  rel/z^2={rel[i]/zz[i]**2:.2e}")
# direct-distance oracle variant (no GEMM trick)
Xa = post.X / post.ls[:, None]; Xb = Xs / post.ls[:, None]
D2 = O._pairwise_sqdist_direct(Xa, Xb)
Ks = post.amp ** 2 * O._kappa(D2, 2)
import scipy.linalg as sl
V = sl.solve_triangular(post.U, Ks, trans='T', lower=False)
var_d = post.amp ** 2 - np.sum(V * V, axis=0) + 1e-18
mu_d = Ks.T @ post.alpha_w
print("direct-distance oracle vs gemm-trick oracle: mu", np.max(np.abs(mu_d - mu_r) / np.abs(mu_r)), "var", np.max(np.abs(var_d - var_r) / var_r))
print("GPU vs direct-distance oracle:               mu", np.max(np.abs(mu - mu_d) / np.abs(mu_d)), "var", np.max(np.abs(var - var_d) / var_d))
