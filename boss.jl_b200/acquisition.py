"""ExpectedImprovement acquisition (mirror of src/acquisitions/expected_improvement.jl and src/acquisition.jl).
The closure returned by construct_acquisition evaluates whole candidate matrices in one library call."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .posterior import model_posterior
from .types import BossOptions, BossProblem, LinFitness, cons_mask


@dataclass
class ExpectedImprovement:
    fitness: object
    eps_samples: int = 200
    cons_safe: bool = True


def best_so_far(problem: BossProblem, fitness):
    """expected_improvement.jl:135-140 (raw observations; feasible = all(y .<= y_max))."""
    Y = problem.data.Y
    if Y.size == 0:
        return None
    feas = np.all(Y <= problem.y_max[:, None], axis=0)
    if not feas.any():
        return None
    return float(max(fitness(Y[:, i]) for i in np.flatnonzero(feas)))


class Acquisition:
    """The safe acquisition closure of construct_safe_acquisition (src/acquisition.jl:21-25):
    acq(x::Vector) -> Real, acq(X::Matrix) -> Vector; -inf where the reference's SafeFunction catches an error,
    0. outside the domain (make_safe, expected_improvement.jl:58-65)."""

    def __init__(self, problem: BossProblem, posteriors, ei: ExpectedImprovement, best):
        if not isinstance(ei.fitness, LinFitness):
            raise NotImplementedError("NonlinFitness (MC-EI over an arbitrary closure) is outside the accelerated path")
        self.problem = problem
        self.posts = posteriors if isinstance(posteriors, list) else [posteriors]
        self.y_dim = problem.data.y_dim
        self.slices = [s.gp for p in self.posts for s in p.slices]
        self.models = [s.model for s in self.posts[0].slices]
        self.coefs = np.asarray(ei.fitness.coefs, dtype=np.float64)
        self.best = best
        self.y_max = None if np.all(np.isinf(problem.y_max)) else problem.y_max
        self.cons_safe = ei.cons_safe

    def device_resident_ok(self):
        """True when nothing on the scoring path is a host closure (prior means, `cons`), i.e. the device-resident
        multi-start driver can run the whole solve."""
        probe = np.zeros((self.problem.data.x_dim, 1))
        no_mean = all(m.mean_at(i, probe) is None for i, m in enumerate(self.models))
        return no_mean and self.problem.domain.cons is None and self.cons_safe

    def _prior_mean(self, X):
        ms = [m.mean_at(i, X) for i, m in enumerate(self.models)]
        if all(v is None for v in ms):
            return None
        return np.stack([np.zeros(X.shape[1]) if v is None else v for v in ms])

    def _guards(self, X):
        if not self.cons_safe:
            return None, None, None
        lb, ub = self.problem.domain.bounds
        return lb, ub, cons_mask(X, self.problem.domain)

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        vec = x.ndim == 1
        X = x[:, None] if vec else x
        lb, ub, cm = self._guards(X)
        acq, _, _ = _lib.ei_score(self.slices, self.y_dim, len(self.posts), X, self.coefs, self.best, self.y_max,
                                  lb, ub, cm, self._prior_mean(X))
        return float(acq[0]) if vec else acq

    def argmax(self, X):
        """-> (index, value) with Julia argmax semantics, one fused launch sequence (no score vector round trip)."""
        X = np.asarray(X, dtype=np.float64)
        lb, ub, cm = self._guards(X)
        _, bv, bi = _lib.ei_score(self.slices, self.y_dim, len(self.posts), X, self.coefs, self.best, self.y_max,
                                  lb, ub, cm, self._prior_mean(X), want_acq=False)
        return int(bi), float(bv)

    def value_and_grad(self, X):
        """-> acq (M,), grad (d, M).  Prior-mean gradients are not propagated for closure means."""
        X = np.asarray(X, dtype=np.float64)
        lb, ub, cm = self._guards(X)
        return _lib.ei_value_grad(self.slices, self.y_dim, len(self.posts), X, self.coefs, self.best, self.y_max,
                                  lb, ub, cm, self._prior_mean(X))


def construct_acquisition(problem: BossProblem, options: BossOptions = BossOptions()) -> Acquisition:
    """construct_acquisition(::ExpectedImprovement, problem, options) (expected_improvement.jl:49-56).
    Refits the posterior exactly like the reference (model_posterior on every call)."""
    ei = problem.acquisition
    post = model_posterior(problem)
    b = best_so_far(problem, ei.fitness)
    if options.info and b is None:
        print("Warning: No feasible solution in the dataset yet. Cannot calculate EI!")
    return Acquisition(problem, post, ei, b)


construct_safe_acquisition = construct_acquisition
