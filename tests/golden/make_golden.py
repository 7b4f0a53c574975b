#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/.

BOSS.jl's own tests hold no numeric golden vectors for the posterior / log-likelihood and the
Julia reference cannot run in the build image, so the fixtures are:

  gp_small_mpmath.json   50-digit mpmath evaluation (direct distances, exact Cholesky) of mean / variance /
                         log marginal likelihood / EI for small problems, all three kernels -- an implementation-
                         independent truth that both the oracle restatement and the CUDA path must reproduce;
  gp_medium_oracle.npz   outputs of oracle/boss_oracle.py (the restated reference path) on a seeded medium problem
                         (n = 200, d = 4): mu, var, EI x PoF, x-gradients, log-likelihood batch -- pins the oracle
                         against silent edits and gives the GPU tests committed numbers;
  reference_known_answers.json   the exact / known-answer cases of the reference's unit tests, re-encoded
                         (test/unit/test/acquisitions/expected_improvement.jl:108-163,
                          test/unit/test/models/gaussian_process.jl:241-259).

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import boss_oracle as O  # noqa: E402
from tests.util_problems import make_hyper_samples, make_problem  # noqa: E402


def mp_ei(mu, var, best):
    import mpmath as mp
    mp.mp.dps = 50
    out = []
    for m, v in zip(mu, var):
        m = mp.mpf(float(m)); s = mp.sqrt(mp.mpf(float(v))); d = m - mp.mpf(float(best))
        z = d / s
        out.append(float(d * mp.ncdf(z) + s * mp.npdf(z)))
    return out


def small_mpmath():
    cases = []
    for kid, n, d, M, seed in ((0, 10, 2, 6, 11), (1, 12, 3, 6, 12), (2, 14, 2, 6, 13), (2, 24, 5, 5, 14)):
        X, Y, ls, amp, ns = make_problem(n, d, seed=seed)
        Xs = np.random.default_rng(seed + 100).random((d, M))
        mu, var, ll = O.adjudicator_mean_var_loglik(X, Y[0], ls[0], amp[0], ns[0], kid, Xs, dps=50)
        best = float(np.quantile(Y[0], 0.75))
        cases.append({"kernel_id": kid, "X": X.tolist(), "y": Y[0].tolist(), "lengthscales": ls[0].tolist(),
                      "amplitude": float(amp[0]), "noise_std": float(ns[0]), "Xs": Xs.tolist(),
                      "mu": mu.tolist(), "var": var.tolist(), "loglik": ll, "best": best,
                      "ei": mp_ei(mu, var, best)})
    with open(os.path.join(HERE, "gp_small_mpmath.json"), "w") as f:
        json.dump({"how": "mpmath dps=50, oracle.adjudicator_mean_var_loglik + closed-form EI", "cases": cases}, f)


def medium_oracle():
    n, d, M, y_dim, S = 200, 4, 64, 2, 8
    X, Y, ls, amp, ns = make_problem(n, d, seed=2024, y_dim=y_dim)
    Xs = np.random.default_rng(2025).random((d, M))
    posts = [O.posterior_fit(X, Y[i], ls[i], amp[i], ns[i], O.KERNEL_MATERN52) for i in range(y_dim)]
    mu = np.stack([O.mean_and_var(p, Xs)[0] for p in posts])
    var = np.stack([O.mean_and_var(p, Xs)[1] for p in posts])
    coefs = np.array([1.0, 0.3])
    y_max = np.array([np.inf, float(np.quantile(Y[1], 0.7))])
    best = O.best_so_far(coefs, Y, y_max)
    acq, _, _ = O.ei_acquisition([posts], Xs, coefs, best, y_max)
    val, grad = O.ei_value_grad(posts, Xs, coefs, best, y_max)
    L, A, N = make_hyper_samples(S, d, seed=2026)
    ll = np.stack([O.gp_loglik_batch(X, Y[0], L, A, N, kid) for kid in (0, 1, 2)])
    cov, _ = O.posterior_cov(posts[0], Xs[:, :16])
    np.savez(os.path.join(HERE, "gp_medium_oracle.npz"), X=X, Y=Y, ls=ls, amp=amp, ns=ns, Xs=Xs, mu=mu, var=var,
             coefs=coefs, y_max=y_max, best=np.array(best), acq=acq, grad=grad, hyp_ls=L, hyp_amp=A, hyp_ns=N,
             loglik=ll, cov16=cov)


def reference_known_answers():
    ka = {
        "clip_var": {"unchanged": [0.0, 1e-9, 1e-8, 1e-7, 1.0], "to_zero": [-1e-9, -1e-8], "domain_error": [-1e-7, -1.0]},
        "expected_improvement": [
            {"coefs": [1.0, 0.0], "mean": [0.0, 0.0], "var": [1.0, 1.0], "best": 0.0, "expect": "positive"},
            {"coefs": [1.0, 0.0], "mean": [0.0, 0.0], "var": [0.0, 0.0], "best": 0.0, "expect": 0.0},
            {"coefs": [1.0, 0.0], "mean": [1.0, 1.0], "var": [0.0, 0.0], "best": 0.0, "expect": 1.0},
            {"coefs": [1.0, 0.0], "mean": [-10.0, -10.0], "var": [1.0, 1.0], "best": 0.0, "expect": "abs<1e-20"}],
        "feas_prob": [
            {"mean": [0.0, 0.0], "var": [1.0, 1.0], "y_max": ["Inf", "Inf"], "expect": 1.0},
            {"mean": [0.0, 0.0], "var": [1.0, 1.0], "y_max": [0.0, "Inf"], "expect": 0.5},
            {"mean": [0.0, 0.0], "var": [1.0, 1.0], "y_max": [0.0, 0.0], "expect": 0.25}],
        "best_so_far": [
            {"coefs": [1.0], "Y": [[1.0, 2.0, 3.0]], "y_max": ["Inf"], "expect": 3.0},
            {"coefs": [1.0], "Y": [[1.0, 2.0, 3.0]], "y_max": [5.0], "expect": 3.0},
            {"coefs": [1.0], "Y": [[10.0, 2.0, 3.0]], "y_max": [5.0], "expect": 3.0},
            {"coefs": [2.0], "Y": [[1.0, 2.0, 3.0]], "y_max": ["Inf"], "expect": 6.0},
            {"coefs": [1.0], "Y": [[1.0, 2.0, 3.0]], "y_max": [0.0], "expect": None}],
    }
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as f:
        json.dump(ka, f, indent=1)


if __name__ == "__main__":
    small_mpmath()
    medium_oracle()
    reference_known_answers()
    print("golden fixtures written to", HERE)
