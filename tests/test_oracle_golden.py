"""Oracle vs the reference's own known-answer tests.

Each case cites the reference test it re-encodes (paths relative to /root/reference).
"""
import math

import numpy as np
import pytest

from oracle import boss_oracle as O


# test/unit/test/models/gaussian_process.jl:241-259
@pytest.mark.parametrize("v", [0.0, 1e-9, 1e-8, 1e-7, 1.0])
def test_clip_var_unchanged(v):
    assert O.clip_var(v) == v


@pytest.mark.parametrize("v", [-1e-9, -1e-8])
def test_clip_var_to_zero(v):
    assert O.clip_var(v) == 0.0


def test_clip_var_domain_error():
    with pytest.raises(O.DomainError):
        O.clip_var(-1e-7)
    with pytest.raises(O.DomainError):
        O.clip_var(float("nan"))


def test_clip_var_status_vectorised():
    v = np.array([0.0, 1e-9, 1.0, -1e-9, -1e-8, -1e-7, np.nan])
    out, st = O.clip_var_status(v)
    assert list(st) == [0, 0, 0, 0, 0, 2, 2]
    assert list(out[:5]) == [0.0, 1e-9, 1.0, 0.0, 0.0]


# test/unit/test/acquisitions/expected_improvement.jl:108-141  (LinFitness([1., 0.]))
COEFS = [1.0, 0.0]


def test_ei_positive():
    assert O.expected_improvement(COEFS, [0.0, 0.0], [1.0, 1.0], 0.0) > 0.0


def test_ei_zero_var_zero_diff():
    assert O.expected_improvement(COEFS, [0.0, 0.0], [0.0, 0.0], 0.0) == 0.0


def test_ei_zero_var_unit_diff():
    assert O.expected_improvement(COEFS, [1.0, 1.0], [0.0, 0.0], 0.0) == 1.0


def test_ei_far_below():
    assert abs(O.expected_improvement(COEFS, [-10.0, -10.0], [1.0, 1.0], 0.0)) <= 1e-20


# test/unit/test/acquisitions/expected_improvement.jl:143-163
def test_feas_prob_cases():
    assert O.feas_prob([0.0, 0.0], [0.0, 0.0], None) == 1.0
    assert O.feas_prob([0.0, 0.0], [1.0, 1.0], None) == 1.0
    assert O.feas_prob([np.inf, np.inf], [1.0, 1.0], None) == 1.0
    assert abs(O.feas_prob([0.0, 0.0], [1.0, 1.0], [np.inf, np.inf]) - 1.0) <= 1e-20
    assert abs(O.feas_prob([0.0, 0.0], [1.0, 1.0], [0.0, np.inf]) - 0.5) <= 1e-20
    assert abs(O.feas_prob([0.0, 0.0], [1.0, 1.0], [0.0, 0.0]) - 0.25) <= 1e-20
    assert 0.99 < O.feas_prob([0.0, 0.0], [1.0, 1.0], [3.0, np.inf]) < 1.0


# test/unit/test/acquisitions/expected_improvement.jl:165-179
def test_best_so_far():
    Y = np.array([[1.0, 2.0, 3.0]])
    assert O.best_so_far([1.0], Y, [np.inf]) == 3.0
    assert O.best_so_far([1.0], Y, [5.0]) == 3.0
    assert O.best_so_far([1.0], np.array([[10.0, 2.0, 3.0]]), [5.0]) == 3.0
    assert O.best_so_far([2.0], Y, [np.inf]) == 6.0
    assert O.best_so_far([1.0], Y, [0.0]) is None
    assert O.best_so_far([1.0], np.zeros((1, 0)), [0.0]) is None


# test/unit/test/acquisitions/expected_improvement.jl:26-40  (make_safe: bounds inclusive, outside -> 0.)
def test_in_bounds_inclusive():
    lb, ub = [5.0], [10.0]
    X = np.array([[1.0, 5.0, 7.0, 10.0, 11.0]])
    assert list(O.in_bounds(X, lb, ub)) == [False, True, True, True, False]


# test/unit/test/utils/utils.jl:33-52
def test_is_feasible():
    assert O.is_feasible([1.0, 2.0], [1.0, 2.0])
    assert O.is_feasible([1.0, 2.0], [np.inf, np.inf])
    assert not O.is_feasible([1.0, 2.1], [1.0, 2.0])


# test/unit/test/models/utils/kernels.jl:13-27  (DiscreteKernel rounds flagged dims)
def test_discrete_kernel_rounding():
    ls = np.array([1.0, 1.0]); mask = np.array([True, False])
    def k(a, b):
        return O.kernel_matrix(np.array(a)[:, None], np.array(b)[:, None], ls, 1.0, O.KERNEL_MATERN32, mask)[0, 0]
    assert k([1.2, 0.5], [0.0, 0.0]) == k([1.0, 0.5], [0.0, 0.0])
    assert k([1.2, 0.5], [0.0, 0.0]) == k([0.8, 0.5], [0.4, 0.0])
    assert k([1.2, 0.5], [0.0, 0.0]) != k([1.2, 0.9], [0.0, 0.0])
    # ties-to-even like Julia's round
    assert k([0.5, 0.0], [0.0, 0.0]) == k([0.0, 0.0], [0.0, 0.0])
    assert k([1.5, 0.0], [2.0, 0.0]) == k([2.0, 0.0], [2.0, 0.0])


def test_julia_argmax_semantics():
    assert O.julia_argmax([1.0, 3.0, 3.0, 2.0]) == 1           # first maximal
    assert O.julia_argmax([1.0, np.nan, 5.0, np.nan]) == 1     # NaN maximal, first NaN
    assert O.julia_argmax([-0.0, 0.0, 0.0]) == 1               # isless(-0.0, 0.0)
    assert O.julia_argmax([-np.inf, -np.inf]) == 0
    rng = np.random.default_rng(0)
    for _ in range(20):
        v = rng.integers(0, 5, 50).astype(float)
        assert O.julia_argmax(v) == O.julia_argmax_fast(v)


def test_normal_cdf_special_cases():
    assert O.normal_cdf(0.0, 0.0, 0.0) == 1.0      # StatsFuns sigma == 0, x == mu
    assert O.normal_cdf(0.0, 0.0, 1.0) == 1.0
    assert O.normal_cdf(0.0, 0.0, -1.0) == 0.0
    assert O.normal_cdf(3.0, 2.0, np.inf) == 1.0   # Infinity() constraint
    assert O.normcdf(0.0) == 0.5


# GP posterior properties the reference tests assert (test/unit/test/models/gaussian_process.jl:57-239)
def _small_post():
    X = np.array([[1.0, 5.0, 9.0], [2.0, 5.0, 8.0]])
    y = np.array([1.0, -1.0, 2.0])
    return X, y, O.posterior_fit(X, y, [2.0, 2.0], 1.0, 1e-3, O.KERNEL_MATERN52)


def test_posterior_vector_matrix_consistency():
    X, y, post = _small_post()
    Xs = np.array([[2.0, 4.0, 7.5], [3.0, 4.0, 1.0]])
    mu, var, st = O.mean_and_var(post, Xs)
    for j in range(Xs.shape[1]):
        m1, v1, _ = O.mean_and_var(post, Xs[:, j])
        assert abs(m1[0] - mu[j]) <= 1e-8 and abs(v1[0] - var[j]) <= 1e-8


def test_posterior_interpolates_and_reverts():
    X, y, post = _small_post()
    mu, var, _ = O.mean_and_var(post, X)
    assert np.allclose(mu, y, atol=0.01)
    mu_far, var_far, _ = O.mean_and_var(post, np.array([[1000.0], [1000.0]]))
    assert abs(mu_far[0]) < 1e-8
    assert abs(var_far[0] - (1.0 + 1e-8) ** 2) < 1e-8
    # variance grows away from data
    _, v, _ = O.mean_and_var(post, np.array([[1.0, 1.5, 2.5], [2.0, 2.5, 3.5]]))
    assert v[0] < v[1] < v[2]


# loglik orderings (test/unit/test/models/gaussian_process.jl:261-316)
def test_loglik_orderings():
    X = np.array([[1.0, 2.0, 3.0, 4.0, 5.0, 6.0]])
    y_alt = np.array([1.0, -1.0, 1.0, -1.0, 1.0, -1.0])
    ll_short = O.gp_loglik(X, y_alt, [0.5], 1.0, 0.1, O.KERNEL_MATERN52)
    ll_long = O.gp_loglik(X, y_alt, [50.0], 1.0, 0.1, O.KERNEL_MATERN52)
    assert ll_short > ll_long
    y = 5.0 * np.sin(X[0])
    assert O.gp_loglik(X, y, [1.0], 1.0, 5.0) > O.gp_loglik(X, y, [1.0], 1.0, 100.0)


def test_loglik_not_pd_is_minus_inf():
    X = np.zeros((1, 3))                 # three identical points, zero noise -> singular K
    assert O.gp_loglik(X, np.array([1.0, 2.0, 3.0]), [1.0], 1.0, 0.0) == -math.inf


def test_mc_expected_improvement_known_answers():
    """expected_improvement(::NonlinFitness, ...) expected_improvement.jl:104-111: with a linear closure and sigma = 0
    the Monte-Carlo estimate is exact, and with symmetric eps = +-1 it is the average of the two branches."""
    fit = lambda y: y[0] + 2.0 * y[1]
    mean = np.array([[1.0, 0.0], [0.5, -1.0]])            # two candidates
    zero = np.zeros((2, 2))
    eps = np.array([[1.0, -1.0], [1.0, -1.0]])
    assert np.array_equal(O.mc_expected_improvement(fit, mean, zero, eps, 1.0), [1.0, 0.0])     # max(0, f - best)
    var = np.array([[1.0, 4.0], [0.25, 1.0]])             # sd = [[1, 2], [0.5, 1]]
    # candidate 0: f = 2 + (1 + 1)*e -> {4, 0} -> improvements {3, 0} -> 1.5 ; candidate 1: f = -2 + 4 e -> {2, -6} -> {1, 0} -> 0.5
    assert np.allclose(O.mc_expected_improvement(fit, mean, var, eps, 1.0), [1.5, 0.5], rtol=0, atol=1e-15)
    # a single eps column (the BI method at :108-111)
    assert np.allclose(O.mc_expected_improvement(fit, mean, var, eps[:, :1], 1.0), [3.0, 1.0], rtol=0, atol=1e-15)
