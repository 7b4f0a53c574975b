"""Fit latency vs rank-1 append latency (host wall clock around the C-ABI call, median of repeats)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import boss_b200
from boss_b200 import _lib
from tests.util_problems import make_problem
_lib.init(0)
for n, d in ((30, 2), (512, 6), (1024, 10), (2048, 8), (4096, 4)):
    X, Y, ls, amp, ns = make_problem(n + 8, d, seed=1)
    tf = []
    for _ in range(5):
        t0 = time.perf_counter(); gp = _lib.gp_fit(X[:, :n], Y[0, :n], ls[0], amp[0], ns[0], 2); tf.append(time.perf_counter() - t0)
        if _ < 4: gp.free()
    ta = []
    for k in range(n, n + 8):
        t0 = time.perf_counter(); _lib.gp_append(gp, X[:, k], Y[0, k]); ta.append(time.perf_counter() - t0)
    gp2 = _lib.gp_fit(X, Y[0], ls[0], amp[0], ns[0], 2)
    print(f"n={n} d={d}: fit {np.median(tf)*1e3:.3f} ms, append {np.median(ta)*1e6:.0f} us (first {ta[0]*1e6:.0f} us), "
          f"loglik diff vs refit {abs(gp.loglik-gp2.loglik)/abs(gp2.loglik):.2e}")
    gp.free(); gp2.free()
