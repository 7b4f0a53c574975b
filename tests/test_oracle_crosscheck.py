"""Corroboration (not a pin) of the unpinned parts of the oracle: posterior mean/var and log
marginal likelihood against scikit-learn's independent implementation and against a 50-digit
mpmath adjudicator; analytic x-gradients against central finite differences."""
import numpy as np
import pytest

from oracle import boss_oracle as O


def _problem(n=40, d=3, seed=7):
    rng = np.random.default_rng(seed)
    X = rng.random((d, n))
    y = np.sin(3 * X).sum(0) + 0.05 * rng.standard_normal(n)
    ls = np.exp(rng.uniform(np.log(0.3), np.log(1.5), d))
    Xs = rng.random((d, 25))
    return X, y, ls, Xs


@pytest.mark.parametrize("kid,nu", [(O.KERNEL_MATERN52, 2.5), (O.KERNEL_MATERN32, 1.5), (O.KERNEL_SE, None)])
def test_against_sklearn(kid, nu):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern
    X, y, ls, Xs = _problem()
    a, s = 1.3, 0.1
    lsc, ac, sc = O.condition_params(ls, a, s)
    base = RBF(length_scale=lsc) if nu is None else Matern(length_scale=lsc, nu=nu)
    gpr = GaussianProcessRegressor(kernel=ConstantKernel(ac * ac) * base, alpha=sc * sc, optimizer=None)
    gpr.fit(X.T, y)
    mu_sk, sd_sk = gpr.predict(Xs.T, return_std=True)
    post = O.posterior_fit(X, y, ls, a, s, kid)
    mu, var, st = O.mean_and_var(post, Xs)
    assert np.all(st == 0)
    assert np.max(np.abs(mu - mu_sk) / np.abs(mu_sk)) < 1e-10
    assert np.max(np.abs(var - sd_sk ** 2) / var) < 1e-9
    ll = O.gp_loglik(X, y, ls, a, s, kid)
    assert abs(ll - gpr.log_marginal_likelihood_value_) / abs(ll) < 1e-12


@pytest.mark.parametrize("kid", [O.KERNEL_SE, O.KERNEL_MATERN32, O.KERNEL_MATERN52])
def test_against_mpmath_adjudicator(kid):
    X, y, ls, Xs = _problem(n=24, d=2, seed=11)
    Xs = Xs[:, :6]
    a, s = 0.9, 0.05
    mu_hp, var_hp, ll_hp = O.adjudicator_mean_var_loglik(X, y, ls, a, s, kid, Xs)
    post = O.posterior_fit(X, y, ls, a, s, kid)
    mu, var = O.mean_and_var_raw(post, Xs)
    assert np.max(np.abs(mu - mu_hp) / np.abs(mu_hp)) < 1e-10
    assert np.max(np.abs(var - var_hp) / np.abs(var_hp)) < 1e-9
    assert abs(O.gp_loglik(X, y, ls, a, s, kid) - ll_hp) / abs(ll_hp) < 1e-12


@pytest.mark.parametrize("kid", [O.KERNEL_SE, O.KERNEL_MATERN32, O.KERNEL_MATERN52])
def test_gradients_vs_finite_differences(kid):
    X, y, ls, Xs = _problem(n=30, d=3, seed=3)
    rng = np.random.default_rng(5)
    Y = np.stack([y, np.cos(2 * X).sum(0) + 0.05 * rng.standard_normal(X.shape[1])])
    posts = [O.posterior_fit(X, Y[i], ls * (1 + 0.2 * i), 1.0, 0.1, kid) for i in range(2)]
    coefs = [1.0, 0.3]; best = float(np.median(np.asarray(coefs) @ Y)); y_max = [np.inf, 1.0]
    val, grad = O.ei_value_grad(posts, Xs, coefs, best, y_max)
    ref, _, _ = O.ei_acquisition([posts], Xs, coefs, best, y_max)
    assert np.allclose(val, ref, rtol=1e-9, atol=0)   # direct vs GEMM-trick distances
    h = 1e-6
    for j in range(X.shape[0]):
        Xp = Xs.copy(); Xp[j] += h
        Xm = Xs.copy(); Xm[j] -= h
        fp, _, _ = O.ei_acquisition([posts], Xp, coefs, best, y_max)
        fm, _, _ = O.ei_acquisition([posts], Xm, coefs, best, y_max)
        fd = (fp - fm) / (2 * h)
        assert np.allclose(grad[j], fd, rtol=2e-5, atol=1e-9), (j, grad[j], fd)


def test_batch_loglik_matches_single():
    X, y, ls, _ = _problem(n=20, d=2)
    rng = np.random.default_rng(1)
    S = 5
    L = np.exp(rng.normal(0, 0.5, (S, 2))); A = np.exp(rng.normal(0, 0.5, S)); N = rng.uniform(0.03, 0.3, S)
    out = O.gp_loglik_batch(X, y, L, A, N, O.KERNEL_SE)
    for s in range(S):
        assert out[s] == O.gp_loglik(X, y, L[s], A[s], N[s], O.KERNEL_SE)


def test_longdouble_adjudicator_matches_mpmath():
    X, y, ls, Xs = _problem(n=24, d=3, seed=5)
    amp, ns = 1.3, 0.07
    for kid in (O.KERNEL_SE, O.KERNEL_MATERN32, O.KERNEL_MATERN52):
        m1, v1, l1 = O.adjudicator_mean_var_loglik(X, y, ls, amp, ns, kid, Xs)
        m2, v2, l2 = O.adjudicator_longdouble(X, y, ls, amp, ns, kid, Xs)
        assert np.max(np.abs(m1 - m2) / np.abs(m1)) <= 1e-14
        assert np.max(np.abs(v1 - v2) / v1) <= 1e-13
        assert abs(l1 - l2) <= 1e-14 * abs(l1)
