"""The BASELINE configs at their FULL sizes, checked through properties that do not need an oracle pass over the whole
input (the oracle scores ~12 k candidates/s): shard / chunk invariance of the argmax, permutation equivariance, equality of
a strided subset with (a) the same points scored alone and (b) the CPU oracle within north_star's tolerances."""
import numpy as np
import pytest

from oracle import boss_oracle as O
from tests.util_problems import make_hyper_samples, make_problem, relerr

pytestmark = pytest.mark.gpu


def test_c2_full_size_two_mi_candidates_per_gpu(lib):
    """configs[1], one GPU's share: n = 2048, d = 8, Matern52, 2 Mi candidates (28 chunks of 592 x 128)."""
    n, d, M = 2048, 8, 1 << 21
    X, Y, ls, amp, ns = make_problem(n, d, seed=1002)
    gp = lib.gp_fit(X, Y[0], ls[0], amp[0], ns[0], lib.KERNEL_MATERN52)
    best = float(np.max(Y[0]))
    Xs = np.random.default_rng(2002).random((M, d)).T
    acq, bv, bi = lib.ei_score([gp], 1, 1, Xs, [1.0], best, None)
    assert acq.shape == (M,) and np.all(np.isfinite(acq)) and np.all(acq >= 0.0)
    assert bi == int(np.argmax(acq)) and bv == acq[bi]                     # first maximal element of the whole batch
    # shards: the argmax of the whole = the best of the per-shard winners (what the 8-GPU run reduces), same bits
    G = 8
    wins = []
    for g in range(G):
        lo, hi = g * (M // G), (g + 1) * (M // G)
        _, v, i = lib.ei_score([gp], 1, 1, Xs[:, lo:hi], [1.0], best, None, want_acq=False)
        wins.append((v, lo + i))
    k = max(range(G), key=lambda g: (wins[g][0], -wins[g][1]))
    assert wins[k] == (bv, bi)
    # permutation: the reversed batch scores to the reversed vector, bit for bit
    acq_r, bv_r, bi_r = lib.ei_score([gp], 1, 1, np.ascontiguousarray(Xs[:, ::-1]), [1.0], best, None)
    assert np.array_equal(acq_r[::-1], acq) and bv_r == bv
    # a strided subset: the same bits when scored alone, the oracle's values to 1e-9, the oracle's argmax
    idx = np.arange(0, M, 4099)
    sub, _, bis = lib.ei_score([gp], 1, 1, np.ascontiguousarray(Xs[:, idx]), [1.0], best, None)
    assert np.array_equal(sub, acq[idx])
    post = O.posterior_fit(X, Y[0], ls[0], amp[0], ns[0], O.KERNEL_MATERN52)
    ref, _, _ = O.ei_acquisition([[post]], Xs[:, idx], [1.0], best, None)
    m = ref > 1e-30 * ref.max()
    assert relerr(sub[m], ref[m]) <= 1e-9
    assert bis == O.julia_argmax_fast(ref)
    gp.free()


def test_c3_full_size_4096_hyperparameter_samples(lib):
    """configs[2] in full: 4096 samples x GP log-likelihood at n = 512, d = 6 (SE and Matern52)."""
    n, d, S = 512, 6, 4096
    X, Y, _, _, _ = make_problem(n, d, seed=1003)
    L, A, N = make_hyper_samples(S, d, seed=3003)
    for kid in (0, 2):
        ll = lib.loglik_batch(X, Y[0], L, A, N, kid)
        assert ll.shape == (S,) and np.all(np.isfinite(ll))
        # shard invariance (8 GPUs x 512 samples) and permutation equivariance, bit for bit
        part = lib.loglik_batch(X, Y[0], L[512:1024], A[512:1024], N[512:1024], kid)
        assert np.array_equal(part, ll[512:1024])
        perm = np.random.default_rng(1).permutation(S)
        assert np.array_equal(lib.loglik_batch(X, Y[0], L[perm], A[perm], N[perm], kid), ll[perm])
        idx = np.arange(0, S, 257)
        ref = O.gp_loglik_batch(X, Y[0], L[idx], A[idx], N[idx], kid)
        assert relerr(ll[idx], ref) <= 1e-8
        assert int(np.argmax(ll[idx])) == int(np.argmax(ref))            # SamplingMAP's pick


def test_c5_full_size_four_outputs_half_mi_candidates(lib):
    """configs[4], one GPU's share: 4 GP outputs with y_max constraints, EI x PoF, n = 1024, d = 10, 512 Ki candidates."""
    n, d, y_dim, M = 1024, 10, 4, 1 << 19
    X, Y, ls, amp, ns = make_problem(n, d, seed=1005, y_dim=y_dim)
    gps = [lib.gp_fit(X, Y[i], ls[i], amp[i], ns[i], lib.KERNEL_MATERN52) for i in range(y_dim)]
    y_max = np.array([np.inf] + [float(np.quantile(Y[i], 0.7)) for i in range(1, y_dim)])
    coefs = np.array([1.0, 0, 0, 0])
    best = O.best_so_far(coefs, Y, y_max)
    lb, ub = np.zeros(d), np.ones(d)
    Xs = np.random.default_rng(5005).random((M, d)).T
    acq, bv, bi = lib.ei_score(gps, y_dim, 1, Xs, coefs, best, y_max, lb=lb, ub=ub)
    assert bi == int(np.argmax(acq)) and bv == acq[bi] and np.all(acq >= 0.0)
    half = M // 2
    _, v0, i0 = lib.ei_score(gps, y_dim, 1, Xs[:, :half], coefs, best, y_max, lb=lb, ub=ub, want_acq=False)
    _, v1, i1 = lib.ei_score(gps, y_dim, 1, Xs[:, half:], coefs, best, y_max, lb=lb, ub=ub, want_acq=False)
    assert (bv, bi) == ((v0, i0) if v0 >= v1 else (v1, half + i1))
    idx = np.arange(0, M, 1031)
    sub, _, _ = lib.ei_score(gps, y_dim, 1, np.ascontiguousarray(Xs[:, idx]), coefs, best, y_max, lb=lb, ub=ub)
    assert np.array_equal(sub, acq[idx])
    posts = [O.posterior_fit(X, Y[i], ls[i], amp[i], ns[i], O.KERNEL_MATERN52) for i in range(y_dim)]
    ref, _, _ = O.ei_acquisition([posts], Xs[:, idx], coefs, best, y_max, lb=lb, ub=ub)
    m = ref > 1e-30 * ref.max()
    assert relerr(sub[m], ref[m]) <= 1e-9
    for g in gps:
        g.free()


def test_c4_full_size_8192_starts(lib):
    """configs[3] in full: n = 4096, d = 4, 8192 multi-start points: value + x-gradient of all starts in one call equal the
    per-shard calls (8 x 1024) bit for bit and the oracle on a subset; the on-device multi-start never loses to its start."""
    n, d, M = 4096, 4, 8192
    X, Y, ls, amp, ns = make_problem(n, d, seed=1004)
    gp = lib.gp_fit(X, Y[0], ls[0], amp[0], ns[0], 2)
    best = float(np.max(Y[0]))
    lb, ub = np.zeros(d), np.ones(d)
    starts = np.random.default_rng(4004).random((d, M))
    a, g = lib.ei_value_grad([gp], 1, 1, starts, [1.0], best, None, lb, ub)
    for s in (0, 5):
        a_s, g_s = lib.ei_value_grad([gp], 1, 1, starts[:, 1024 * s:1024 * (s + 1)], [1.0], best, None, lb, ub)
        assert np.array_equal(a_s, a[1024 * s:1024 * (s + 1)]) and np.array_equal(g_s, g[:, 1024 * s:1024 * (s + 1)])
    idx = np.arange(0, M, 257)
    post = O.posterior_fit(X, Y[0], ls[0], amp[0], ns[0], 2)
    a_r, g_r = O.ei_value_grad([post], starts[:, idx], [1.0], best, None)
    m = a_r > 1e-30 * a_r.max()
    assert relerr(a[idx][m], a_r[m]) <= 1e-9
    assert np.max(np.abs(g[:, idx][:, m] - g_r[:, m]) / np.max(np.abs(g_r[:, m]), axis=0)) <= 1e-7
    Xo, fo, bx, bv, bi, _ = lib.ei_maximize_multistart([gp], 1, 1, starts[:, :1024], [1.0], best, None, lb, ub, iters=20)
    assert np.all(fo >= a[:1024] - 1e-15) and bv == fo[bi] and bi == int(np.argmax(fo))
    gp.free()
