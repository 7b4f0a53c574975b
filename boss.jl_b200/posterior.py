"""model_posterior dispatch (mirror of src/posterior.jl): problem -> params -> slices, broadcast over BI samples."""
from __future__ import annotations

import numpy as np

from .gaussian_process import model_posterior_slice
from .types import BossProblem, get_params


class DefaultModelPosterior:
    """src/posterior.jl:31-79: stacks y_dim independent slices."""

    def __init__(self, slices):
        self.slices = slices

    def mean_and_var(self, x):
        res = [s.mean_and_var(x) for s in self.slices]
        x = np.asarray(x)
        if x.ndim == 1:
            return np.array([r[0] for r in res]), np.array([r[1] for r in res])
        return np.stack([r[0] for r in res]), np.stack([r[1] for r in res])      # y_dim x M

    def mean(self, x):
        return self.mean_and_var(x)[0]

    def var(self, x):
        return self.mean_and_var(x)[1]

    def std(self, x):
        return np.sqrt(self.var(x))


def model_posterior(problem_or_model, params=None, data=None):
    """model_posterior(problem) | model_posterior(model, params, data); a list of params (BIParams)
    broadcasts to a list of posteriors (src/posterior.jl:15-16)."""
    if isinstance(problem_or_model, BossProblem):
        prob = problem_or_model
        return model_posterior(prob.model, get_params(prob.params), prob.data)
    model = problem_or_model
    params = get_params(params)
    if isinstance(params, (list, tuple)):
        return [model_posterior(model, p, data) for p in params]
    return DefaultModelPosterior([model_posterior_slice(model, params, data, i) for i in range(data.y_dim)])


def average_mean(posteriors, x):
    """src/posterior.jl:177-179"""
    return sum(p.mean(x) for p in posteriors) / len(posteriors)
