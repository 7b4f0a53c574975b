"""GPU tests of the reference-facing host mirror (boss_b200.*) -- they read like the reference's own tests
(test/unit/test/models/gaussian_process.jl, test/unit/test/acquisitions/expected_improvement.jl)."""
import numpy as np
import pytest

import boss_b200 as B
from oracle import boss_oracle as O

pytestmark = pytest.mark.gpu


def _problem(noise=0.1, y_max=None, discrete=None, cons=None, seed=0, n=20, f=None):
    rng = np.random.default_rng(seed)
    f = f or (lambda x: np.array([np.sin(x[0]) + 0.5 * np.cos(2 * x[1])]))
    dom = B.Domain(([0.0, 0.0], [10.0, 10.0]), discrete=discrete, cons=cons)
    X = B.generate_LHC(dom.bounds, n, rng)
    if discrete is not None:
        X = np.array([B.types.discrete_round(discrete, X[:, j]) for j in range(n)]).T
    Y = np.stack([f(X[:, j]) for j in range(n)], axis=1)
    model = B.GaussianProcess(
        kernel=B.SqExponentialKernel(),
        lengthscale_priors=[B.mvlognormal([0.5, 0.5], [0.5, 0.5])] * Y.shape[0],
        amplitude_priors=[B.LogNormal(0.0, 0.5)] * Y.shape[0],
        noise_std_priors=[B.Dirac(noise)] * Y.shape[0])
    return B.BossProblem(f, dom, B.ExpectedImprovement(B.LinFitness([1.0] + [0.0] * (Y.shape[0] - 1))), model,
                         B.ExperimentData(X, Y), y_max=y_max)


@pytest.fixture(scope="module", autouse=True)
def _init(lib):
    return lib


def test_gp_posterior_properties():
    # test/unit/test/models/gaussian_process.jl:57-126: 3-point 2-D data
    X = np.array([[1.0, 5.0, 9.0], [2.0, 5.0, 8.0]]); Y = np.array([[1.0, -1.0, 2.0]])
    model = B.GaussianProcess(kernel=B.Matern52Kernel())
    params = B.GaussianProcessParams(np.array([[2.0], [2.0]]), np.array([1.0]), np.array([1e-3]))
    post = B.model_posterior(model, params, B.ExperimentData(X, Y))
    Xs = np.array([[2.0, 4.0, 7.5], [3.0, 4.0, 1.0]])
    mus, vs = post.mean_and_var(Xs)
    assert mus.shape == (1, 3) and vs.shape == (1, 3)
    for j in range(3):
        m1, v1 = post.mean_and_var(Xs[:, j])
        assert abs(m1[0] - mus[0, j]) <= 1e-8 and abs(v1[0] - vs[0, j]) <= 1e-8
    assert np.allclose(post.mean(X)[0], Y[0], atol=0.01)
    assert abs(post.mean(np.array([1000.0, 1000.0]))[0]) < 1e-8
    v = post.var(np.array([[1.0, 1.5, 2.5], [2.0, 2.5, 3.5]]))[0]
    assert v[0] < v[1] < v[2]


def test_construct_acquisition_and_gridam_match_oracle():
    prob = _problem(seed=3)
    prob.params = B.estimate_parameters(B.SamplingMAP(64, seed=1), prob)
    acq = B.construct_acquisition(prob)
    assert isinstance(acq(np.array([1.0, 1.0])), float)
    assert acq(np.array([11.0, 1.0])) == 0.0                         # out of bounds -> 0.
    am = B.GridAM(prob, steps=[0.25, 0.25], shuffle=False)
    x, val = B.maximize_acquisition(am, prob)
    p = prob.params.params
    post = O.posterior_fit(prob.data.X, prob.data.Y[0], p.lengthscales[:, 0], p.amplitudes[0], p.noise_std[0], O.KERNEL_SE)
    ref, _, _ = O.ei_acquisition([[post]], am.points, [1.0], B.best_so_far(prob, prob.acquisition.fitness), None,
                                 *prob.domain.bounds)
    k = O.julia_argmax_fast(ref)
    assert np.array_equal(x, am.points[:, k]) and abs(val - ref[k]) <= 1e-9 * abs(ref[k])


def test_sampling_map_picks_first_best_of_the_batch():
    prob = _problem(seed=4)
    fit = B.SamplingMAP(50, seed=7)
    res = B.estimate_parameters(fit, prob)
    # replay the same prior draws through the oracle
    sampler = prob.model.params_sampler(np.random.default_rng(7))
    samples = [sampler() for _ in range(50)]
    pll = prob.model.params_loglike()
    vals = np.array([O.gp_loglik(prob.data.X, prob.data.Y[0], s.lengthscales[:, 0], s.amplitudes[0], s.noise_std[0],
                                 O.KERNEL_SE) + pll(s) for s in samples])
    b = int(np.argmax(vals))
    assert np.array_equal(res.params.lengthscales, samples[b].lengthscales)
    assert abs(res.loglike - vals[b]) <= 1e-8 * abs(vals[b])


def test_optimization_am_beats_grid_and_respects_bounds():
    prob = _problem(seed=5)
    prob.params = B.estimate_parameters(B.SamplingMAP(64, seed=2), prob)
    _, gval = B.maximize_acquisition(B.GridAM(prob, steps=[0.5, 0.5], shuffle=False), prob)
    x, val = B.maximize_acquisition(B.OptimizationAM(multistart=64, iters=40, seed=0), prob)
    assert B.in_bounds(x, prob.domain.bounds)
    assert val >= gval * (1 - 1e-6)


def test_constrained_discrete_multi_output_problem_runs():
    f = lambda x: np.array([np.sin(x[0]) + 0.1 * x[1], np.cos(x[0]) - 0.05 * x[1]])
    prob = _problem(y_max=[np.inf, 0.5], discrete=[True, False], cons=lambda x: [x[0] + x[1] - 1.0], seed=6, f=f)
    prob.params = B.estimate_parameters(B.SamplingMAP(32, seed=3), prob)
    acq = B.construct_acquisition(prob)
    vals = acq(np.array([[0.2, 3.0, 3.4], [0.2, 3.0, 3.0]]))
    assert vals[0] == 0.0                                            # cons violated
    assert vals[1] == vals[2]                                        # DiscreteKernel: 3.4 rounds to 3.0
    x, val = B.maximize_acquisition(B.SamplingAM(None, 500, seed=1), prob)
    assert B.in_domain(x, prob.domain)


def test_bo_loop_example_style():
    """BASELINE configs[0] style: 2-D SE-ARD GP, 20 initial points, OptimizationMAP + OptimizationAM, EI."""
    prob = _problem(seed=8)
    n0 = len(prob.data)
    B.bo(prob, B.OptimizationMAP(multistart=8, iters=10, seed=0), B.OptimizationAM(multistart=32, iters=20, seed=0),
         B.IterLimit(3))
    assert len(prob.data) == n0 + 3 and prob.consistent
    assert np.isfinite(prob.params.loglike)


def test_semiparametric_loglike_uses_per_sample_mean():
    rng = np.random.default_rng(9)
    X = rng.random((2, 40)) * 4; Y = (2.0 * X[0] + np.sin(X[1]))[None, :]
    gp = B.GaussianProcess(kernel=B.Matern52Kernel(), lengthscale_priors=[B.mvlognormal([0, 0], [0.3, 0.3])],
                           amplitude_priors=[B.LogNormal()], noise_std_priors=[B.Dirac(0.1)])
    model = B.Semiparametric(B.Parametric(lambda x, th: np.array([th[0] * x[0] + th[1]]), [B.Uniform(-3, 3), B.Uniform(-1, 1)]), gp)
    data = B.ExperimentData(X, Y)
    ll = B.data_loglike(model, data)
    mk = lambda th: B.SemiparametricParams(np.array(th), np.array([[1.0], [1.0]]), np.array([1.0]), np.array([0.1]))
    good, bad = ll(mk([2.0, 0.0])), ll(mk([-2.0, 0.0]))
    assert good > bad                                                # loglik ordering incl. theta (semiparametric tests)
    both = ll([mk([2.0, 0.0]), mk([-2.0, 0.0])])
    assert both[0] == good and both[1] == bad
    ref = O.gp_loglik(X, Y[0] - 2.0 * X[0], [1.0, 1.0], 1.0, 0.1, O.KERNEL_MATERN52)
    assert abs(good - ref) <= 1e-8 * abs(ref)


def test_sequential_batch_am_matches_refit_per_pick():
    """SequentialBatchAM (batch.jl:26-38) with the appended factor cache == the reference's refit-per-pick loop."""
    prob = _problem(seed=11, n=25)
    prob.params = B.estimate_parameters(B.SamplingMAP(48, seed=2), prob)
    inner = B.GridAM(prob, steps=[0.5, 0.5], shuffle=False)
    Xb, _ = B.maximize_acquisition(B.SequentialBatchAM(inner, 4), prob)
    assert Xb.shape == (2, 4) and len(prob.data) == 25                 # the caller's problem is not modified
    # reference loop: refit from scratch after every speculative point
    ref = B.BossProblem(prob.f, prob.domain, prob.acquisition, prob.model,
                        B.ExperimentData(prob.data.X.copy(), prob.data.Y.copy()), params=prob.params)
    picks = []
    for _ in range(4):
        post = B.model_posterior(ref)
        x, _ = B.maximize_acquisition(inner, ref)
        ref.data.augment(x, post.mean(x))
        picks.append(x)
    assert np.array_equal(Xb, np.stack(picks, axis=1))
    assert len({tuple(c) for c in Xb.T}) == 4                          # speculative points push the picks apart


def test_data_loglike_and_grad_matches_finite_differences():
    prob = _problem(seed=5, n=18, f=lambda x: np.array([np.sin(x[0]), np.cos(x[1]) + 0.1 * x[0]]))
    rng = np.random.default_rng(0)
    p = B.GaussianProcessParams(rng.uniform(1.0, 3.0, (2, 2)), rng.uniform(0.5, 1.5, 2), rng.uniform(0.05, 0.2, 2))
    ll = B.data_loglike(prob.model, prob.data)
    llg = B.data_loglike_and_grad(prob.model, prob.data)
    v, g = llg(p)
    assert abs(v - ll(p)) <= 1e-9 * abs(v)
    h = 1e-6
    for name in ("lengthscales", "amplitudes", "noise_std"):
        arr = getattr(p, name)
        for idx in np.ndindex(arr.shape):
            pp = B.GaussianProcessParams(p.lengthscales.copy(), p.amplitudes.copy(), p.noise_std.copy())
            pm = B.GaussianProcessParams(p.lengthscales.copy(), p.amplitudes.copy(), p.noise_std.copy())
            getattr(pp, name)[idx] += h
            getattr(pm, name)[idx] -= h
            fd = (ll(pp) - ll(pm)) / (2 * h)
            assert abs(getattr(g, name)[idx] - fd) <= 1e-5 * max(1.0, abs(fd)), (name, idx)
    vals, grads = llg([p, p])
    assert vals[0] == vals[1] == v and np.array_equal(grads[1].lengthscales, g.lengthscales)


def test_nonlin_fitness_mc_ei_matches_oracle_map_and_bi():
    """NonlinFitness (expected_improvement.jl:104-111): posterior mean / variance from the device, Monte-Carlo average
    over the same eps on the host; single posterior (all eps columns) and BI posteriors (one column each), with
    y_max constraints, bounds and Domain.cons guards."""
    f = lambda x: np.array([np.sin(x[0]) + 0.1 * x[1], np.cos(x[0]) - 0.05 * x[1]])
    fit_scalar = lambda y: float(np.cos(y[0]) + np.sin(y[1]))                 # the reference's docstring example
    fit_vector = lambda y: np.cos(y[0]) + np.sin(y[1])                        # same, broadcasting over columns
    Xs = np.random.default_rng(1).random((2, 257)) * 11.0 - 0.5               # some points out of bounds
    for fit in (fit_scalar, fit_vector):
        prob = _problem(y_max=[np.inf, 0.5], cons=lambda x: [x[0] + x[1] - 1.0], seed=6, f=f)
        prob.acquisition = B.ExpectedImprovement(B.NonlinFitness(fit), eps_samples=37)
        prob.params = B.estimate_parameters(B.SamplingMAP(32, seed=3), prob)
        eps = np.random.default_rng(5).standard_normal((2, 37))
        acq = B.construct_acquisition(prob, eps=eps)
        p = prob.params.params
        posts = [[O.posterior_fit(prob.data.X, prob.data.Y[i], p.lengthscales[:, i], p.amplitudes[i], p.noise_std[i],
                                  O.KERNEL_SE) for i in range(2)]]
        best = B.best_so_far(prob, prob.acquisition.fitness)
        assert best is not None
        cm = B.types.cons_mask(Xs, prob.domain)
        ref = O.mc_ei_acquisition(posts, Xs, fit_scalar, eps, best, prob.y_max, *prob.domain.bounds, cons_mask=cm)
        got = acq(Xs)
        assert np.any(ref > 0) and np.any(ref == 0.0)
        assert np.all(np.abs(got - ref) <= 1e-9 * np.maximum(np.abs(ref), 1e-300))
        assert acq(Xs[:, 3]) == got[3]
        k, v = acq.argmax(Xs)
        assert k == O.julia_argmax_fast(ref) and v == got[k]
        with pytest.raises(NotImplementedError):
            acq.value_and_grad(Xs)
    # BI: three hyper-parameter samples, eps has one column per posterior
    prob.params = B.estimate_parameters(B.SamplingMAP(32, seed=3), prob)
    plist = []
    for s in range(3):
        q = prob.params.params
        plist.append(B.GaussianProcessParams(q.lengthscales * (1.0 + 0.2 * s), q.amplitudes * (1.0 + 0.1 * s), q.noise_std))
    post_bi = B.model_posterior(prob.model, plist, prob.data)
    eps3 = np.random.default_rng(7).standard_normal((2, 3))
    acq_bi = B.acquisition.Acquisition(prob, post_bi, prob.acquisition, best, eps3)
    posts = [[O.posterior_fit(prob.data.X, prob.data.Y[i], q.lengthscales[:, i], q.amplitudes[i], q.noise_std[i], O.KERNEL_SE)
              for i in range(2)] for q in plist]
    ref = O.mc_ei_acquisition(posts, Xs, fit_scalar, eps3, best, prob.y_max, *prob.domain.bounds, cons_mask=cm)
    got = acq_bi(Xs)
    assert np.all(np.abs(got - ref) <= 1e-9 * np.maximum(np.abs(ref), 1e-300))


def test_expr_fitness_runs_mc_ei_on_the_device_and_matches_the_host_closure_path():
    """ExprFitness = a NonlinFitness from the device's expression set: ExpectedImprovement with it evaluates the
    Monte-Carlo EI (expected_improvement.jl:104-111) inside boss_mcei_score; the same fitness wrapped as an opaque
    NonlinFitness closure goes through the host finish -- both must agree, and agree with the oracle."""
    f = lambda x: np.array([np.sin(x[0]) + 0.1 * x[1], np.cos(x[0]) - 0.05 * x[1]])
    Xs = np.random.default_rng(11).random((2, 300)) * 11.0 - 0.5
    eps = np.random.default_rng(12).standard_normal((2, 64))
    for ef in (B.ExprFitness("quadratic", [1.0, 0.3], c0=0.1, q=[0.0, -0.5], t=[0.0, 0.2]),
               B.ExprFitness("min", [1.0, 2.0], t=[0.0, 0.5]), B.ExprFitness("affine", [0.7, -0.2], c0=1.0)):
        prob = _problem(y_max=[np.inf, 0.6], cons=lambda x: [x[0] + x[1] - 1.0], seed=6, f=f)
        prob.acquisition = B.ExpectedImprovement(ef, eps_samples=64)
        prob.params = B.estimate_parameters(B.SamplingMAP(32, seed=3), prob)
        acq_dev = B.construct_acquisition(prob, eps=eps)
        assert acq_dev.expr is not None
        prob2 = _problem(y_max=[np.inf, 0.6], cons=lambda x: [x[0] + x[1] - 1.0], seed=6, f=f)
        prob2.acquisition = B.ExpectedImprovement(B.NonlinFitness(lambda y, ef=ef: ef(y)), eps_samples=64)
        prob2.params = prob.params
        acq_host = B.construct_acquisition(prob2, eps=eps)
        assert acq_host.expr is None
        a, b = acq_dev(Xs), acq_host(Xs)
        assert np.any(b > 0)
        assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(b))
        k, v = acq_dev.argmax(Xs)
        assert k == int(np.argmax(a)) and v == a[k]


def test_gradient_optimization_map_and_sample_opt_map():
    """OptimizationMAP with the analytic hyper-parameter gradient (boss_gp_loglik_grad_batch; the reference's default
    autodiff path, optimization.jl:41,153) and SampleOptMAP (sample_opt.jl:37-48): both must reach at least the
    log-likelihood of the best of their starts, and the gradient search must match or beat the compass search."""
    prob = _problem(seed=8, n=40)
    ll = B.model_loglike(prob.model, prob.data)
    samp = B.estimate_parameters(B.SamplingMAP(64, seed=21), prob)
    grad = B.estimate_parameters(B.OptimizationMAP(multistart=12, iters=30, seed=21, algorithm="lbfgs"), prob)
    comp = B.estimate_parameters(B.OptimizationMAP(multistart=12, iters=30, seed=21, algorithm="compass"), prob)
    so = B.estimate_parameters(B.SampleOptMAP(samples=64, multistart=6, iters=30, seed=21), prob)
    for r in (grad, comp, so):
        assert abs(ll(r.params) - r.loglike) <= 1e-8 * abs(r.loglike)          # reported value is the value at the point
        assert np.all(r.params.noise_std == 0.1)                                 # Dirac-pinned noise stays exact
    assert so.loglike >= samp.loglike - 1e-9                                     # starts from the best samples
    assert grad.loglike >= comp.loglike - 0.05 * abs(comp.loglike)
    allr = B.estimate_parameters(B.OptimizationMAP(multistart=5, iters=5, seed=3), prob, return_all=True)
    assert 1 <= len(allr) <= 5


def test_model_posterior_testset_of_the_reference():
    """test/unit/test/models/gaussian_process.jl:57-160 ("model_posterior(model, params, data)"), assertion by assertion:
    two outputs, prior mean x -> [1, 1], data on the diagonal (y = x at 2, 5, 8), noise 1e-4, parameters estimated by
    SamplingMAP(samples = 200); vector / matrix / single-column-matrix forms of mean, std, var, cov and their mean_and_*
    pairs agree to 1e-8, the posterior interpolates the data to 0.01, reverts to the prior mean far away, and the
    variance grows away from the data."""
    X = np.array([[2.0, 5.0, 8.0], [2.0, 5.0, 8.0]])
    model = B.GaussianProcess(mean=lambda x: [1.0, 1.0],
                              amplitude_priors=[B.LogNormal(0.0, 1.0)] * 2,
                              lengthscale_priors=[B.mvlognormal([1.0, 1.0], [1.0, 1.0])] * 2,
                              noise_std_priors=[B.Dirac(1e-4)] * 2)
    problem = B.BossProblem(lambda x: x, B.Domain(([0.0, 0.0], [10.0, 10.0])),
                            B.ExpectedImprovement(B.LinFitness([1.0, 0.0])), model, B.ExperimentData(X, X.copy()),
                            y_max=[np.inf, 5.0])
    problem.params = B.estimate_parameters(B.SamplingMAP(samples=200, seed=5), problem)
    out = B.model_posterior(problem.model, problem.params, problem.data)
    x2 = np.array([2.0, 2.0])
    # vector
    for f in (out.mean, out.std, out.var):
        assert np.asarray(f(x2)).shape == (2,)
    assert np.allclose(out.mean(x2), out.mean_and_std(x2)[0], atol=1e-8)
    assert np.allclose(out.mean(x2), out.mean_and_var(x2)[0], atol=1e-8)
    assert np.allclose(out.std(x2), out.mean_and_std(x2)[1], atol=1e-8)
    assert np.allclose(out.var(x2), out.mean_and_var(x2)[1], atol=1e-8)
    for v in (2.0, 5.0, 8.0):
        assert np.allclose(out.mean(np.array([v, v])), [v, v], atol=0.01)
    assert np.all(out.mean(np.array([1.0, 1.0])) < [2.0, 2.0])
    assert np.all(out.mean(np.array([4.0, 4.0])) < [5.0, 5.0])
    assert np.allclose(out.mean(np.array([100.0, 100.0])), [1.0, 1.0], atol=0.01)
    assert np.all(out.var(x2) <= out.var(np.array([3.0, 3.0])))
    assert np.all(out.var(np.array([10.0, 10.0])) <= out.var(np.array([11.0, 11.0])))
    # matrix, and single-element matrix
    for Xm in (np.array([[1.0, 2.0, 3.0], [1.0, 2.0, 3.0]]), np.array([[1.0], [1.0]])):
        M = Xm.shape[1]
        assert out.mean(Xm).shape == (2, M) and out.std(Xm).shape == (2, M) and out.var(Xm).shape == (2, M)
        assert out.cov(Xm).shape == (M, M, 2)
        assert np.allclose(out.mean(Xm), out.mean_and_std(Xm)[0], atol=1e-8)
        assert np.allclose(out.mean(Xm), out.mean_and_var(Xm)[0], atol=1e-8)
        assert np.allclose(out.mean(Xm), out.mean_and_cov(Xm)[0], atol=1e-8)
        assert np.allclose(out.std(Xm), out.mean_and_std(Xm)[1], atol=1e-8)
        assert np.allclose(out.var(Xm), out.mean_and_var(Xm)[1], atol=1e-8)
        assert np.allclose(out.cov(Xm), out.mean_and_cov(Xm)[1], atol=1e-8)
        for j in range(M):
            assert np.allclose(out.mean(Xm)[:, j], out.mean(Xm[:, j]), atol=1e-8)
            assert np.allclose(out.var(Xm)[:, j], out.var(Xm[:, j]), atol=1e-8)
            # the diagonal of the covariance is the variance (both clipped by _clip_var)
            assert np.allclose(out.cov(Xm)[j, j, :], out.var(Xm[:, j]), atol=1e-8)
    # slices: model_posterior_slice agrees with the stacked posterior (src/posterior.jl:4-5, gaussian_process.jl:133-141)
    for i in range(2):
        sl = B.model_posterior_slice(problem.model, problem.params.params if hasattr(problem.params, "params") else problem.params,
                                     problem.data, i)
        assert abs(sl.mean(x2) - out.mean(x2)[i]) <= 1e-8 and abs(sl.var(x2) - out.var(x2)[i]) <= 1e-8
        assert abs(sl.std(x2) - out.std(x2)[i]) <= 1e-8


def test_loglike_testsets_of_the_reference():
    """test/unit/test/models/gaussian_process.jl:244-356: the "model_loglike", "data_loglike" and "params_loglike"
    testsets with the reference's own inputs (3 training points -> the warp-register path; noise_std = 0 included)."""
    GPP = B.GaussianProcessParams
    ones22 = np.ones((2, 2))
    # --- model_loglike(model, data) ---
    model = B.GaussianProcess(mean=lambda x: [1.0, 1.0], amplitude_priors=[B.LogNormal()] * 2,
                              lengthscale_priors=[B.mvlognormal([1.0, 1.0], [1.0, 1.0])] * 2,
                              noise_std_priors=[B.LogNormal()] * 2)
    X = np.array([[2.0, 5.0, 8.0], [2.0, 5.0, 8.0]])
    out = B.model_loglike(model, B.ExperimentData(X, X.copy()))
    assert callable(out)
    v = out(GPP(ones22, np.array([1.0, 1.0]), np.array([1.0, 1.0])))
    assert isinstance(v, float) and v < 0.0
    assert out(GPP(ones22, np.ones(2), np.array([5.0, 5.0]))) > out(GPP(ones22, np.ones(2), np.array([100.0, 100.0])))
    assert out(GPP(ones22, np.array([5.0, 5.0]), np.ones(2))) > out(GPP(ones22, np.array([100.0, 100.0]), np.ones(2)))
    assert out(GPP(5.0 * ones22, np.ones(2), np.ones(2))) > out(GPP(100.0 * ones22, np.ones(2), np.ones(2)))
    # --- data_loglike(model, data) ---
    model1 = B.GaussianProcess(mean=lambda x: [0.0], amplitude_priors=[B.LogNormal()],
                               lengthscale_priors=[B.Product([B.Dirac(1.0)])], noise_std_priors=[B.Dirac(0.1)])
    X1 = np.array([[1.0, 2.0, 3.0]])
    t_ls = lambda l: GPP(np.array([[l]]), np.array([1.0]), np.array([0.0]))
    t_amp = lambda a: GPP(np.array([[1.0]]), np.array([a]), np.array([0.0]))
    t_ns = lambda s: GPP(np.array([[1.0]]), np.array([1.0]), np.array([s]))
    assert isinstance(B.data_loglike(model1, B.ExperimentData(X1, X1.copy()))(GPP(np.array([[1.0]]), np.array([1.0]), np.array([1.0]))), float)
    out = B.data_loglike(model1, B.ExperimentData(X1, np.array([[1.0, -1.0, 1.0]])))
    assert out(t_ls(0.1)) > out(t_ls(1.0)) > out(t_ls(10.0))
    assert out(t_amp(1.0)) > out(t_amp(0.1))
    assert out(t_ns(1.0)) > out(t_ns(0.1)) and out(t_ns(1.0)) > out(t_ns(10.0))
    t_data = lambda Y: B.data_loglike(model1, B.ExperimentData(X1, np.array([Y])))(GPP(np.array([[1.0]]), np.array([1.0]), np.array([0.0])))
    assert t_data([99.9, 100.0, 100.1]) > t_data([99.0, 100.0, 101.0]) > t_data([90.0, 100.0, 110.0])
    # the device values are the oracle's (the reference's logpdf(::FiniteGP)) to 1e-8
    for p in (t_ls(0.1), t_ls(10.0), t_amp(0.1), t_ns(10.0)):
        ref = O.gp_loglik(X1, np.array([1.0, -1.0, 1.0]), p.lengthscales[:, 0], p.amplitudes[0], p.noise_std[0], O.KERNEL_MATERN52)
        assert abs(out(p) - ref) <= 1e-8 * abs(ref)
    # --- params_loglike(model, params) ---
    m = B.GaussianProcess(lengthscale_priors=[B.mvlognormal([1.0, 1.0], [1.0, 1.0])] * 2, amplitude_priors=[B.LogNormal()] * 2,
                          noise_std_priors=[B.Dirac(0.1)] * 2)
    assert isinstance(m.params_loglike()(GPP(ones22, np.array([1.0, 2.0]), np.array([0.1, 0.1]))), float)
    md = B.GaussianProcess(lengthscale_priors=[B.Product([B.Dirac(1.0), B.Dirac(1.0)])] * 2, amplitude_priors=[B.Dirac(1.0)] * 2,
                           noise_std_priors=[B.Dirac(0.1)] * 2)
    pl = md.params_loglike()
    assert pl(GPP(ones22, np.ones(2), np.array([0.1, 0.1]))) == 0.0
    assert pl(GPP(np.array([[1.0, 5.0], [1.0, 5.0]]), np.ones(2), np.array([0.1, 0.1]))) == -np.inf
    assert pl(GPP(ones22, np.array([1.0, 5.0]), np.array([0.1, 0.1]))) == -np.inf
    assert pl(GPP(ones22, np.ones(2), np.array([0.1, 0.5]))) == -np.inf


def test_semiparametric_model_posterior_testset_of_the_reference():
    """test/unit/test/models/semiparametric.jl:1-130: Semiparametric(NonlinearModel + GP residual) on the reference's
    own problem (theta ~ Normal^4, noise 1e-4, SamplingMAP(samples = 200)); vector / matrix forms agree, the variance
    grows away from the data, and the GP residual interpolates the data on top of the parametric mean."""
    X = np.array([[2.0, 5.0, 8.0], [2.0, 5.0, 8.0]])
    fY = lambda x: np.array([np.sin(x[0]) + np.exp(x[1]), np.cos(x[0]) + np.exp(x[1])])
    Y = np.stack([fY(X[:, j]) for j in range(3)], axis=1)
    predict = lambda x, th: np.array([th[0] * np.sin(x[0]) + th[1] * np.exp(x[1]), th[2] * np.cos(x[0]) + th[3] * np.exp(x[1])])
    model = B.Semiparametric(B.Parametric(predict, [B.Normal()] * 4),
                             B.GaussianProcess(amplitude_priors=[B.LogNormal()] * 2,
                                               lengthscale_priors=[B.mvlognormal([1.0, 1.0], [1.0, 1.0])] * 2,
                                               noise_std_priors=[B.Dirac(1e-4)] * 2))
    problem = B.BossProblem(lambda x: x, B.Domain(([0.0, 0.0], [10.0, 10.0])), B.ExpectedImprovement(B.LinFitness([1.0, 0.0])),
                            model, B.ExperimentData(X, Y), y_max=[np.inf, 5.0])
    problem.params = B.estimate_parameters(B.SamplingMAP(samples=200, seed=6), problem)
    out = B.model_posterior(problem.model, problem.params, problem.data)
    x2 = np.array([2.0, 2.0])
    assert np.asarray(out.mean(x2)).shape == (2,) and np.asarray(out.std(x2)).shape == (2,) and np.asarray(out.var(x2)).shape == (2,)
    assert np.allclose(out.mean(x2), out.mean_and_std(x2)[0], atol=1e-8) and np.allclose(out.mean(x2), out.mean_and_var(x2)[0], atol=1e-8)
    assert np.allclose(out.std(x2), out.mean_and_std(x2)[1], atol=1e-8) and np.allclose(out.var(x2), out.mean_and_var(x2)[1], atol=1e-8)
    assert np.all(out.var(x2) <= out.var(np.array([3.0, 3.0])))
    assert np.all(out.var(np.array([10.0, 10.0])) <= out.var(np.array([11.0, 11.0])))
    Xm = np.array([[1.0, 2.0, 3.0], [1.0, 2.0, 3.0]])
    assert out.mean(Xm).shape == (2, 3) and out.std(Xm).shape == (2, 3) and out.var(Xm).shape == (2, 3) and out.cov(Xm).shape == (3, 3, 2)
    assert np.allclose(out.mean(Xm), out.mean_and_cov(Xm)[0], atol=1e-8) and np.allclose(out.cov(Xm), out.mean_and_cov(Xm)[1], atol=1e-8)
    assert np.allclose(out.var(Xm), out.mean_and_var(Xm)[1], atol=1e-8) and np.allclose(out.std(Xm), out.mean_and_std(Xm)[1], atol=1e-8)
    for j in range(3):
        assert np.allclose(out.mean(Xm)[:, j], out.mean(Xm[:, j]), atol=1e-8) and np.allclose(out.var(Xm)[:, j], out.var(Xm[:, j]), atol=1e-8)
    # the data are interpolated (noise 1e-4) whatever theta was drawn: parametric mean + GP residual.  exp(8) = 2981, so
    # "to 0.01" is relative to the data scale here
    for j in range(3):
        assert np.allclose(out.mean(X[:, j]), Y[:, j], atol=0.01 + 1e-5 * np.max(np.abs(Y[:, j])))


def test_average_mean_testset_of_the_reference():
    """test/unit/test/posterior.jl: average_mean over BI posteriors (8 hyper-parameter samples -- the reference draws them
    with TuringBI, out of scope here: 8 prior samples stand in; one boss_gp_fit_batch call per output slice fits them)."""
    X = np.array([[2.0, 5.0, 8.0], [2.0, 5.0, 8.0]])
    model = B.GaussianProcess(amplitude_priors=[B.LogNormal()] * 2, lengthscale_priors=[B.mvlognormal([1.0, 1.0], [1.0, 1.0])] * 2,
                              noise_std_priors=[B.Dirac(1e-4)] * 2)
    problem = B.BossProblem(lambda x: x, B.Domain(([0.0, 0.0], [10.0, 10.0])), B.ExpectedImprovement(B.LinFitness([1.0, 0.0])),
                            model, B.ExperimentData(X, X.copy()), y_max=[np.inf, 5.0])
    sampler = model.params_sampler(np.random.default_rng(12))
    problem.params = B.BIParams([sampler() for _ in range(8)])
    posts = B.model_posterior(problem)
    assert len(posts) == 8
    for x in (np.array([3.0, 3.0]), np.array([5.0, 5.0])):
        assert np.allclose(B.average_mean(posts, x), sum(p.mean(x) for p in posts) / len(posts), rtol=1e-12, atol=0)
    # the batched fit is the one-by-one fit, bit for bit
    one = B.model_posterior(problem.model, problem.params.samples[3], problem.data)
    assert np.array_equal(one.mean(np.array([3.0, 3.0])), posts[3].mean(np.array([3.0, 3.0])))
    assert np.array_equal(one.var(np.array([3.0, 3.0])), posts[3].var(np.array([3.0, 3.0])))
