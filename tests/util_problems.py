"""Seeded synthetic problems shared by the GPU parity tests and bench.py."""
import numpy as np


def make_problem(n, d, seed, y_dim=1, noise=0.1):
    rng = np.random.default_rng(seed)
    X = rng.random((d, n))
    Y = np.empty((y_dim, n))
    for i in range(y_dim):
        Y[i] = np.sin(3 * X + 0.7 * i).sum(0) + 0.05 * rng.standard_normal(n)
    ls = np.exp(rng.uniform(np.log(0.3), np.log(1.5), (y_dim, d)))
    amp = np.full(y_dim, 1.0)
    ns = np.full(y_dim, noise)
    return X, Y, ls, amp, ns


def make_hyper_samples(S, d, seed):
    rng = np.random.default_rng(seed)
    ls = np.exp(rng.normal(0.0, 0.5, (S, d)))
    amp = np.exp(rng.normal(0.0, 0.5, S))
    ns = np.exp(rng.uniform(np.log(0.03), np.log(0.3), S))
    return ls, amp, ns


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0
