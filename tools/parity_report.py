#!/usr/bin/env python
"""profiles/r02_parity_report.json: max relative error of every quantity of the hot path, CUDA (through the C ABI) vs the
CPU oracle, on the shape of every BASELINE config (candidate / sample counts bounded so that the oracle finishes in
seconds).  Tolerances are north_star's: 1e-9 for posterior mean / variance / EI, 1e-8 for the log marginal likelihood,
argmax index exact.  Run on a GPU box:  python tools/parity_report.py > gpurun_out/parity_report.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import boss_b200  # noqa: F401
    from boss_b200 import _lib
    from oracle import boss_oracle as O
    from tests.util_problems import make_hyper_samples, make_problem, relerr
    _lib.init(0)
    rows = []

    def scoring(tag, n, d, kid, M, y_dim=1, seed=0, cons=False, S_ll=16, grad_pts=64):
        t0 = time.time()
        X, Y, ls, amp, ns = make_problem(n, d, seed=seed, y_dim=y_dim)
        gps = [_lib.gp_fit(X, Y[i], ls[i], amp[i], ns[i], kid) for i in range(y_dim)]
        posts = [O.posterior_fit(X, Y[i], ls[i], amp[i], ns[i], kid) for i in range(y_dim)]
        Xs = np.random.default_rng(seed + 1).random((d, M))
        y_max = np.array([np.inf] + [float(np.quantile(Y[i], 0.7)) for i in range(1, y_dim)]) if cons else None
        coefs = np.zeros(y_dim); coefs[0] = 1.0
        best = O.best_so_far(coefs, Y, y_max) if cons else float(np.max(Y[0]))
        lb, ub = np.zeros(d), np.ones(d)
        row = {"config": tag, "n": n, "d": d, "kernel": ["SE", "Matern32", "Matern52"][kid], "y_dim": y_dim, "candidates": M}
        e_mu = e_var = 0.0
        for i in range(y_dim):
            mu, var, _ = _lib.gp_predict(gps[i], Xs)
            mo, vo, _ = O.mean_and_var(posts[i], Xs)
            e_mu = max(e_mu, float(np.max(np.abs(mu - mo)) / np.max(np.abs(mo))))
            e_var = max(e_var, relerr(var, vo))
        row["mean_max_abs_err_rel_to_max"] = e_mu
        row["var_max_relerr"] = e_var
        acq, bv, bi = _lib.ei_score(gps, y_dim, 1, Xs, coefs, best, y_max, lb=lb, ub=ub)
        ref, _, _ = O.ei_acquisition([posts], Xs, coefs, best, y_max, lb=lb, ub=ub)
        m = ref > 1e-30 * ref.max()        # deeper in the tail EI's relative error is z^2-amplified (tests/test_gpu_parity.py)
        row["ei_max_relerr"] = relerr(acq[m], ref[m])
        row["ei_points_compared"] = int(m.sum())
        row["argmax_index_equal"] = bool(bi == O.julia_argmax_fast(ref))
        k = min(grad_pts, M)
        a_g, g_g = _lib.ei_value_grad(gps, y_dim, 1, Xs[:, :k], coefs, best, y_max, lb=lb, ub=ub)
        a_r, g_r = O.ei_value_grad(posts, Xs[:, :k], coefs, best, y_max)
        mm = a_r > 1e-30 * a_r.max()
        if mm.any():
            row["ei_grad_max_err_rel_to_grad_norm"] = float(np.max(np.abs(g_g[:, mm] - g_r[:, mm]) / np.max(np.abs(g_r[:, mm]), axis=0)))
        L, A, N = make_hyper_samples(S_ll, d, seed=seed + 2)
        ll = _lib.loglik_batch(X, Y[0], L, A, N, kid)
        ll_ref = O.gp_loglik_batch(X, Y[0], L, A, N, kid)
        row["loglik_max_relerr"] = relerr(ll, ll_ref)
        row["loglik_samples"] = S_ll
        if n <= 1024:
            llg, gr = _lib.loglik_grad_batch(X, Y[0], L[:8], A[:8], N[:8], kid)
            llr, grr = O.gp_loglik_grad_batch(X, Y[0], L[:8], A[:8], N[:8], kid)
            row["loglik_grad_max_err_rel_to_grad_norm"] = float(np.max(np.abs(gr - grr) / np.linalg.norm(grr, axis=1, keepdims=True)))
        row["seconds"] = round(time.time() - t0, 1)
        for g in gps:
            g.free()
        rows.append(row)

    scoring("C1 example.jl style (n=20, d=2, SE)", 20, 2, 0, 2000, seed=101, S_ll=64)
    scoring("C2 headline (n=2048, d=8, Matern52)", 2048, 8, 2, 16384, seed=1002, S_ll=8)
    scoring("C3 TuringBI batch (n=512, d=6, Matern52)", 512, 6, 2, 4096, seed=1003, S_ll=64)
    scoring("C3 TuringBI batch (n=512, d=6, SE)", 512, 6, 0, 4096, seed=1003, S_ll=64)
    scoring("C4 Semiparametric residual GP (n=4096, d=4, Matern52)", 4096, 4, 2, 2048, seed=1004, S_ll=2)
    scoring("C5 constrained 4-output (n=1024, d=10, Matern52)", 1024, 10, 2, 8192, y_dim=4, seed=1005, cons=True, S_ll=8)
    tol = {"mean_max_abs_err_rel_to_max": 1e-9, "var_max_relerr": 1e-9, "ei_max_relerr": 1e-9, "loglik_max_relerr": 1e-8,
           "ei_grad_max_err_rel_to_grad_norm": 1e-7, "loglik_grad_max_err_rel_to_grad_norm": 1e-8}
    worst = {k: max((r[k] for r in rows if k in r), default=None) for k in tol}
    ok = all(r["argmax_index_equal"] for r in rows) and all(worst[k] is None or worst[k] <= tol[k] for k in tol)
    print(json.dumps({"what": "CUDA path (C ABI) vs CPU oracle (oracle/boss_oracle.py) on every BASELINE config shape",
                      "tolerances": tol, "worst": worst, "all_within_tolerance": bool(ok),
                      "oracle_pin": "partial: the reference's own known answers pin _clip_var / EI closed form / feas_prob / "
                                    "best_so_far; posterior mean / variance / LML values are corroborated by scikit-learn and a "
                                    "50-digit mpmath adjudicator (tests/test_oracle_crosscheck.py), not by Julia outputs",
                      "rows": rows}, indent=1))
    _lib.shutdown()


if __name__ == "__main__":
    main()
