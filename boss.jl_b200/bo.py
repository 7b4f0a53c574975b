"""bo! main loop (mirror of src/bo.jl:30-59) -- orchestration only; kept so that example-style scripts run."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .acquisition_maximizers import maximize_acquisition
from .model_fitters import estimate_parameters
from .types import BossOptions, BossProblem


@dataclass
class IterLimit:
    iter_max: int
    iter: int = 0

    def __call__(self, problem):
        if self.iter >= self.iter_max:
            return False
        self.iter += 1
        return True


def bo(problem: BossProblem, model_fitter, acq_maximizer, term_cond=None, options: BossOptions = BossOptions()):
    term_cond = term_cond or IterLimit(1)
    if not problem.consistent:
        problem.params = estimate_parameters(model_fitter, problem, options)          # bo.jl:80-84
        problem.consistent = True
    while term_cond(problem):
        x, _ = maximize_acquisition(acq_maximizer, problem, options)                   # bo.jl:95
        y = np.atleast_1d(problem.f(x))                                                # bo.jl:117
        problem.data.augment(x[:, None], y[:, None])
        problem.consistent = False
        problem.params = estimate_parameters(model_fitter, problem, options)
        problem.consistent = True
        if options.callback:
            options.callback(problem)
    return problem
