"""Edge cases of the C ABI on the device: empty and ragged batches, maximum sizes, NaN / Inf semantics,
argument errors (the reference's behaviour is cited per case)."""
import numpy as np
import pytest

from oracle import boss_oracle as O
from tests.util_problems import make_hyper_samples, make_problem, relerr

pytestmark = pytest.mark.gpu


def test_empty_batches(lib):
    X, Y, ls, amp, ns = make_problem(12, 2, seed=1)
    gp = lib.gp_fit(X, Y[0], ls[0], amp[0], ns[0], 2)
    acq, bv, bi = lib.ei_score([gp], 1, 1, np.empty((2, 0)), [1.0], 0.0, None)
    assert acq.shape == (0,) and bi == -1                   # argmax of an empty collection: no index
    mu, var, st = lib.gp_predict(gp, np.empty((2, 0)))
    assert mu.shape == var.shape == st.shape == (0,)
    out = lib.loglik_batch(X, Y[0], np.empty((0, 2)), np.empty(0), np.empty(0), 2)
    assert out.shape == (0,)
    gp.free()


@pytest.mark.parametrize("M", [1, 127, 128, 129, 255, 257, 75775, 75777])
def test_ragged_candidate_counts(lib, M):
    """Batch sizes around the 128-candidate block and the 592-block chunk boundaries."""
    n, d = 64, 2
    X, Y, ls, amp, ns = make_problem(n, d, seed=2)
    gp = lib.gp_fit(X, Y[0], ls[0], amp[0], ns[0], 0)
    post = O.posterior_fit(X, Y[0], ls[0], amp[0], ns[0], 0)
    Xs = np.random.default_rng(M).random((d, M))
    best = float(np.median(Y[0]))
    acq, bv, bi = lib.ei_score([gp], 1, 1, Xs, [1.0], best, None)
    ref, _, _ = O.ei_acquisition([[post]], Xs, [1.0], best, None)
    m = ref > 1e-200
    assert relerr(acq[m], ref[m]) <= 1e-9 and bi == O.julia_argmax_fast(ref) and bv == acq[bi]
    gp.free()


def test_single_training_point_and_max_dims(lib):
    # n = 1
    gp = lib.gp_fit(np.array([[0.3], [0.7]]), np.array([1.5]), [0.5, 0.5], 1.0, 0.1, 2)
    post = O.posterior_fit(np.array([[0.3], [0.7]]), np.array([1.5]), [0.5, 0.5], 1.0, 0.1, 2)
    Xs = np.random.default_rng(0).random((2, 50))
    mu, var, _ = lib.gp_predict(gp, Xs)
    mr, vr, _ = O.mean_and_var(post, Xs)
    assert relerr(mu, mr) <= 1e-9 and relerr(var, vr) <= 1e-9
    gp.free()
    # x_dim = 32 is the largest supported dimension, 33 is an argument error; y_dim = 16 likewise
    n, d = 40, 32
    X, Y, ls, amp, ns = make_problem(n, d, seed=3, y_dim=16)
    gps = [lib.gp_fit(X, Y[i], ls[i] * 6, amp[i], ns[i], 1) for i in range(16)]
    posts = [O.posterior_fit(X, Y[i], ls[i] * 6, amp[i], ns[i], 1) for i in range(16)]
    Xs = np.random.default_rng(1).random((d, 300))
    coefs = np.linspace(1.0, 0.1, 16)
    y_max = np.where(np.arange(16) % 2 == 0, np.inf, 3.0)
    best = O.best_so_far(coefs, Y, y_max)
    acq, _, bi = lib.ei_score(gps, 16, 1, Xs, coefs, best, y_max)
    ref, _, _ = O.ei_acquisition([posts], Xs, coefs, best, y_max)
    m = ref > 1e-200
    assert relerr(acq[m], ref[m]) <= 1e-9 and bi == O.julia_argmax_fast(ref)
    with pytest.raises(lib.BossError):
        lib.gp_fit(np.zeros((33, 5)), np.zeros(5), np.ones(33), 1.0, 0.1, 2)
    with pytest.raises(lib.BossError):
        lib.ei_score(gps + gps[:1], 17, 1, Xs, np.ones(17), 0.0, None)
    for g in gps:
        g.free()


def test_nan_candidate_is_maximal_like_julia_argmax(lib):
    """Julia's argmax treats NaN as maximal.  A NaN coordinate makes the posterior NaN, the
    variance check fails (NaN is not >= -1e-8) and the reference's SafeFunction returns -Inf -- so the NaN never
    reaches argmax; an out-of-bounds NaN is caught by the bounds guard first (comparison false -> 0.)."""
    X, Y, ls, amp, ns = make_problem(30, 2, seed=4)
    gp = lib.gp_fit(X, Y[0], ls[0], amp[0], ns[0], 2)
    Xs = np.random.default_rng(2).random((2, 10))
    Xs[0, 3] = np.nan
    acq, bv, bi = lib.ei_score([gp], 1, 1, Xs, [1.0], float(np.median(Y[0])), None)
    assert acq[3] == -np.inf and bi != 3 and np.isfinite(bv)
    mu, var, st = lib.gp_predict(gp, Xs)
    assert st[3] == 2 and st.sum() == 2                      # only that candidate carries the DomainError status
    gp.free()


def test_infinite_constraint_and_zero_variance_paths(lib):
    """cdf(., Infinity()) == 1 exactly (src/utils/inf.jl:13-15); y_max == mu with sigma == 0 -> cdf 1 (StatsFuns)."""
    X, Y, ls, amp, ns = make_problem(25, 2, seed=5, y_dim=2)
    gps = [lib.gp_fit(X, Y[i], ls[i], amp[i], ns[i], 0) for i in range(2)]
    Xs = np.random.default_rng(3).random((2, 200))
    a1, _, _ = lib.ei_score(gps, 2, 1, Xs, [1.0, 0.0], 0.1, [np.inf, np.inf])
    a2, _, _ = lib.ei_score(gps, 2, 1, Xs, [1.0, 0.0], 0.1, None)
    assert np.array_equal(a1, a2)
    for g in gps:
        g.free()


def test_argument_errors_do_not_poison_the_library(lib):
    X, Y, ls, amp, ns = make_problem(20, 2, seed=6)
    with pytest.raises(lib.BossError):
        lib.gp_fit(X, Y[0], ls[0], -1.0, ns[0], 2)           # assert alpha >= 0 (gaussian_process.jl:228)
    with pytest.raises(lib.BossError):
        lib.gp_fit(X, Y[0], ls[0], amp[0], ns[0], 7)         # unknown kernel
    gp = lib.gp_fit(X, Y[0], ls[0], amp[0], ns[0], 2)        # the library still works afterwards
    with pytest.raises(lib.BossError):
        lib.ei_score([gp], 1, 1, np.zeros((2, 3)), [1.0], 0.0, None, lb=np.zeros(2))      # lb without ub
    acq, _, _ = lib.ei_score([gp], 1, 1, np.full((2, 3), 0.5), [1.0], 0.0, None)
    assert np.all(np.isfinite(acq))
    gp.free()
    gp.free()                                                # double free is a no-op
    with pytest.raises(lib.BossError):
        lib.gp_predict(gp, np.zeros((2, 1)))                 # use after free is an error, not a crash


def test_large_training_set_n4096_semiparametric_style(lib):
    """BASELINE config C4 shape: n = 4096, d = 4, prior mean from a parametric model, value + gradient."""
    n, d, M = 4096, 4, 300
    X, Y, ls, amp, ns = make_problem(n, d, seed=1004)
    theta = np.array([0.8, -0.3])
    mean_X = theta[0] * X[0] + theta[1]
    y = Y[0] + mean_X
    gp = lib.gp_fit(X, y - mean_X, ls[0], amp[0], ns[0], 2)
    post = O.posterior_fit(X, y - mean_X, ls[0], amp[0], ns[0], 2)
    Xs = np.random.default_rng(7).random((d, M))
    pm = (theta[0] * Xs[0] + theta[1])[None, :]
    pmg = np.zeros((1, d, M)); pmg[0, 0] = theta[0]
    best = float(np.quantile(y, 0.9))
    val, grad = lib.ei_value_grad([gp], 1, 1, Xs, [1.0], best, None, prior_mean_s=pm, prior_mean_grad_s=pmg)
    ref, gref = O.ei_value_grad([post], Xs, [1.0], best, None, prior_mean_s=pm, prior_mean_grad_s=pmg)
    mu, var, _ = lib.gp_predict(gp, Xs, pm[0])
    mr, vr, _ = O.mean_and_var(post, Xs, pm[0])
    assert relerr(mu, mr) <= 1e-9 and relerr(var, vr) <= 1e-9
    m = ref > 1e-30 * ref.max()                              # deeper in the tail EI's relative error is z^2-amplified
    assert relerr(val[m], ref[m]) <= 1e-9
    assert np.max(np.abs(grad[:, m] - gref[:, m]) / np.max(np.abs(gref[:, m]), axis=0)) <= 1e-8
    gp.free()


def test_concurrent_callers_are_serialised(lib):
    """The reference may call the acquisition / log-likelihood from several Julia tasks at once when parallel=true
    (Threads.@threads in src/utils/optim_multistart.jl:62).  Entry points serialise on an internal mutex: results
    from 4 concurrent host threads equal the serial results bit for bit."""
    import threading
    n, d = 300, 3
    X, Y, ls, amp, ns = make_problem(n, d, seed=9)
    gp = lib.gp_fit(X, Y[0], ls[0], amp[0], ns[0], 2)
    rng = np.random.default_rng(10)
    Xs = [rng.random((d, 3000 + 500 * t)) for t in range(4)]
    L, A, N = make_hyper_samples(12, d, seed=11)
    best = float(np.median(Y[0]))
    serial = [(lib.ei_score([gp], 1, 1, x, [1.0], best, None)[0], lib.loglik_batch(X, Y[0], L, A, N, 2)) for x in Xs]
    out = [None] * 4

    def work(t):
        for _ in range(3):
            out[t] = (lib.ei_score([gp], 1, 1, Xs[t], [1.0], best, None)[0], lib.loglik_batch(X, Y[0], L, A, N, 2))
    ths = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for t in range(4):
        assert np.array_equal(out[t][0], serial[t][0]) and np.array_equal(out[t][1], serial[t][1])
    gp.free()


def test_large_training_set_n_6100(lib):
    """48 row blocks (n = 6100 is not a multiple of 128): right-looking fit + triangular inverse, scoring chunks and
    the left-looking batched log-likelihood far beyond the benchmark shapes."""
    n, d = 6100, 5
    X, Y, ls, amp, ns = make_problem(n, d, seed=n)
    gp = lib.gp_fit(X, Y[0], ls[0], amp[0], ns[0], lib.KERNEL_MATERN52)
    post = O.posterior_fit(X, Y[0], ls[0], amp[0], ns[0], O.KERNEL_MATERN52)
    Xs = np.random.default_rng(1).random((d, 2000))
    mu, var, st = lib.gp_predict(gp, Xs)
    mu_r, var_r, _ = O.mean_and_var(post, Xs)
    assert relerr(mu, mu_r) <= 1e-9 and relerr(var, var_r) <= 1e-9 and not st.any()
    ll_r = O.gp_loglik(X, Y[0], ls[0], amp[0], ns[0], O.KERNEL_MATERN52)
    ll = lib.loglik_batch(X, Y[0], np.vstack([ls[0], 1.3 * ls[0]]), np.r_[amp[0], amp[0]], np.r_[ns[0], ns[0]],
                          lib.KERNEL_MATERN52)
    assert abs(gp.loglik - ll_r) <= 1e-8 * abs(ll_r) and abs(ll[0] - ll_r) <= 1e-8 * abs(ll_r)
    gp.free()
