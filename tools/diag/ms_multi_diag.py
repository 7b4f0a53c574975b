"""Diagnose multi-device vs single-device multi-start differences (run on >= 2 GPUs)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import boss_b200  # noqa
from boss_b200 import _lib as lib
from oracle import boss_oracle as O
from tests.util_problems import make_problem
import torch
nd = int(os.environ.get("ND", torch.cuda.device_count()))
n, d = 300, 4
X, Y, ls, amp, ns = make_problem(n, d, seed=101, y_dim=2)
rng = np.random.default_rng(102)
_ = rng.random((d, 150_000))
y_max = np.array([np.inf, float(np.quantile(Y[1], 0.7))])
coefs = np.array([1.0, 0.0])
best = O.best_so_far(coefs, Y, y_max)
starts = rng.random((d, 512 * nd))
lib.init(0)
gps = [lib.gp_fit(X, Y[i], ls[i], amp[i], ns[i], 2) for i in range(2)]
ref = lib.ei_maximize_multistart(gps, 2, 1, starts, coefs, best, y_max, np.zeros(d), np.ones(d), iters=15)
# single device, shard by shard (same library path as the multi-device call, sequentially)
xs, fs = [], []
for k in range(nd):
    o = lib.ei_maximize_multistart(gps, 2, 1, starts[:, 512 * k:512 * (k + 1)], coefs, best, y_max, np.zeros(d), np.ones(d), iters=15)
    xs.append(o[0]); fs.append(o[1])
xs = np.concatenate(xs, axis=1); fs = np.concatenate(fs)
bad = np.where(fs != ref[1])[0]
print("single-device shard-by-shard vs whole: differing starts", len(bad), "max |df|", float(np.max(np.abs(fs - ref[1]))), "first", bad[:8])
for g in gps:
    g.free()
lib.init_multi(nd)
gps = [lib.gp_fit(X, Y[i], ls[i], amp[i], ns[i], 2) for i in range(2)]
o2 = lib.ei_maximize_multistart(gps, 2, 1, starts, coefs, best, y_max, np.zeros(d), np.ones(d), iters=15)
bad = np.where(o2[1] != ref[1])[0]
print("multi-device vs whole: differing starts", len(bad), "max |df|", float(np.max(np.abs(o2[1] - ref[1]))), "first", bad[:8])
bad = np.where(o2[1] != fs)[0]
print("multi-device vs shard-by-shard: differing starts", len(bad))
