"""Committed golden fixtures (tests/golden/, generator: tests/golden/make_golden.py).

CPU part: the oracle reproduces them (pins the oracle).  GPU part: the CUDA path, through the C ABI,
reproduces them within the north-star tolerances (1e-9 posterior / EI, 1e-8 log-likelihood)."""
import json
import os

import numpy as np
import pytest

from oracle import boss_oracle as O
from tests.util_problems import relerr

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_POST, TOL_LL = 1e-9, 1e-8


def _small():
    with open(os.path.join(G, "gp_small_mpmath.json")) as f:
        return json.load(f)["cases"]


def _medium():
    return np.load(os.path.join(G, "gp_medium_oracle.npz"))


def _inf(v):
    return [np.inf if x == "Inf" else x for x in v]


# ------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("i", range(4))
def test_oracle_matches_mpmath_golden(i):
    c = _small()[i]
    X, Xs = np.array(c["X"]), np.array(c["Xs"])
    post = O.posterior_fit(X, c["y"], c["lengthscales"], c["amplitude"], c["noise_std"], c["kernel_id"])
    mu, var, _ = O.mean_and_var(post, Xs)
    assert relerr(mu, c["mu"]) <= 1e-10 and relerr(var, c["var"]) <= 1e-10
    ll = O.gp_loglik(X, c["y"], c["lengthscales"], c["amplitude"], c["noise_std"], c["kernel_id"])
    assert abs(ll - c["loglik"]) <= 1e-11 * abs(c["loglik"])
    acq, _, _ = O.ei_acquisition([[post]], Xs, [1.0], c["best"], None)
    assert relerr(acq, c["ei"]) <= 1e-9


def test_oracle_matches_medium_golden():
    g = _medium()
    posts = [O.posterior_fit(g["X"], g["Y"][i], g["ls"][i], g["amp"][i], g["ns"][i], O.KERNEL_MATERN52) for i in range(2)]
    for i in range(2):
        mu, var, _ = O.mean_and_var(posts[i], g["Xs"])
        assert relerr(mu, g["mu"][i]) <= 1e-12 and relerr(var, g["var"][i]) <= 1e-11
    acq, _, _ = O.ei_acquisition([posts], g["Xs"], g["coefs"], float(g["best"]), g["y_max"])
    assert relerr(acq, g["acq"]) <= 1e-10
    for k, kid in enumerate((0, 1, 2)):
        ll = O.gp_loglik_batch(g["X"], g["Y"][0], g["hyp_ls"], g["hyp_amp"], g["hyp_ns"], kid)
        assert relerr(ll, g["loglik"][k]) <= 1e-12


def test_reference_known_answers_on_oracle():
    with open(os.path.join(G, "reference_known_answers.json")) as f:
        ka = json.load(f)
    for v in ka["clip_var"]["unchanged"]:
        assert O.clip_var(v) == v
    for v in ka["clip_var"]["to_zero"]:
        assert O.clip_var(v) == 0.0
    for v in ka["clip_var"]["domain_error"]:
        with pytest.raises(O.DomainError):
            O.clip_var(v)
    for c in ka["expected_improvement"]:
        out = float(O.expected_improvement(c["coefs"], np.array(c["mean"])[:, None], np.array(c["var"])[:, None], c["best"])[0])
        if c["expect"] == "positive":
            assert out > 0.0
        elif c["expect"] == "abs<1e-20":
            assert abs(out) < 1e-20
        else:
            assert out == c["expect"]
    for c in ka["feas_prob"]:
        out = float(O.feas_prob(np.array(c["mean"])[:, None], np.array(c["var"])[:, None], np.array(_inf(c["y_max"])))[0])
        assert abs(out - c["expect"]) <= 1e-20
    for c in ka["best_so_far"]:
        assert O.best_so_far(np.array(c["coefs"]), np.array(c["Y"]), np.array(_inf(c["y_max"]))) == c["expect"]


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("i", range(4))
def test_cuda_matches_mpmath_golden(lib, i):
    c = _small()[i]
    X, Xs = np.array(c["X"]), np.array(c["Xs"])
    gp = lib.gp_fit(X, c["y"], c["lengthscales"], c["amplitude"], c["noise_std"], c["kernel_id"])
    mu, var, st = lib.gp_predict(gp, Xs)
    assert relerr(mu, c["mu"]) <= TOL_POST and relerr(var, c["var"]) <= TOL_POST and not st.any()
    assert abs(gp.loglik - c["loglik"]) <= TOL_LL * abs(c["loglik"])
    ll = lib.loglik_batch(X, c["y"], np.array([c["lengthscales"]]), np.array([c["amplitude"]]), np.array([c["noise_std"]]),
                          c["kernel_id"])
    assert abs(ll[0] - c["loglik"]) <= TOL_LL * abs(c["loglik"])          # warp-register path (n <= 32)
    acq, _, bi = lib.ei_score([gp], 1, 1, Xs, [1.0], c["best"], None)
    assert relerr(acq, c["ei"]) <= TOL_POST
    assert bi == int(np.argmax(c["ei"]))
    gp.free()


@pytest.mark.gpu
def test_cuda_matches_medium_golden(lib):
    g = _medium()
    gps = [lib.gp_fit(g["X"], g["Y"][i], g["ls"][i], g["amp"][i], g["ns"][i], lib.KERNEL_MATERN52) for i in range(2)]
    for i in range(2):
        mu, var, _ = lib.gp_predict(gps[i], g["Xs"])
        assert relerr(mu, g["mu"][i]) <= TOL_POST and relerr(var, g["var"][i]) <= TOL_POST
    acq, bv, bi = lib.ei_score(gps, 2, 1, g["Xs"], g["coefs"], float(g["best"]), g["y_max"])
    assert relerr(acq, g["acq"]) <= TOL_POST and bi == int(np.argmax(g["acq"]))
    val, grad = lib.ei_value_grad(gps, 2, 1, g["Xs"], g["coefs"], float(g["best"]), g["y_max"])
    assert np.max(np.abs(grad - g["grad"]) / np.max(np.abs(g["grad"]), axis=1, keepdims=True)) <= 1e-8
    for k, kid in enumerate((0, 1, 2)):
        ll = lib.loglik_batch(g["X"], g["Y"][0], g["hyp_ls"], g["hyp_amp"], g["hyp_ns"], kid)
        assert relerr(ll, g["loglik"][k]) <= TOL_LL
    _, cov, _ = lib.gp_cov(gps[0], g["Xs"][:, :16])
    sc = np.sqrt(np.outer(np.diag(g["cov16"]), np.diag(g["cov16"])))
    assert np.max(np.abs(cov - g["cov16"]) / sc) <= TOL_POST
    for gp in gps:
        gp.free()
