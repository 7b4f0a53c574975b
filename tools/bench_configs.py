#!/usr/bin/env python
"""Per-config numbers for BASELINE.json configs C3 / C4 / C5: per-GPU share of each workload,
device-resident inputs, CUDA events on the library stream, plus a parity spot check against the oracle on a
subset.  One JSON line per config.  (bench.py carries C2 + the headline log-likelihood shape.)

  python tools/bench_configs.py [--configs c3,c4,c5] [--steps 3] [--full]   (--full: whole-job sizes on one GPU)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


_PEAK = [None]
_DIST = [None]     # torch.distributed when every rank runs its share (times are then the max over ranks)


def peak():
    """FP64 denominator: the DGEMM rate bench.py measured in this run if it set one, else the round-1 file."""
    if _PEAK[0] is not None:
        return _PEAK[0]
    with open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")) as f:
        return float(json.load(f)["fp64_tflops_peak_used"])


def _max_over_ranks(torch, ms):
    d = _DIST[0]
    if d is None:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def timed(torch, stream, fn, steps, warmup=3, warm_ms=150.0):
    """>= `warmup` untimed calls and >= `warm_ms` of untimed work (SM clocks drop within milliseconds of idling --
    e.g. while the CPU oracle computes the parity subset -- and a few short calls do not bring them back)."""
    t0 = time.perf_counter(); k = 0
    while k < warmup or (time.perf_counter() - t0) * 1e3 < warm_ms:
        fn(); k += 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return _max_over_ranks(torch, e0.elapsed_time(e1) / steps)


def f_k(d):
    return 3 * d + 8


def run_c3(torch, _lib, O, stream, args):
    from tests.util_problems import make_hyper_samples, make_problem, relerr
    n, d = 512, 6
    S = 4096 if args.full else 512
    X, Y, _, _, _ = make_problem(n, d, seed=1003)
    out = []
    for kid, kname in ((2, "Matern52"), (0, "SE")):
        L, A, N = make_hyper_samples(S, d, seed=3003)
        tX = torch.tensor(np.ascontiguousarray(X.T), device="cuda"); ty = torch.tensor(Y[0], device="cuda")
        tL = torch.tensor(L, device="cuda"); tA = torch.tensor(A, device="cuda"); tN = torch.tensor(N, device="cuda")
        tll = torch.empty(S, dtype=torch.float64, device="cuda")

        def step():
            _lib.loglik_batch_dev(tX.data_ptr(), d, n, ty.data_ptr(), 0, tL.data_ptr(), tA.data_ptr(), tN.data_ptr(), kid, S,
                                  tll.data_ptr())
        ms = timed(torch, stream, step, args.steps)
        ll = tll.cpu().numpy()
        k = min(S, 48)
        ref = O.gp_loglik_batch(X, Y[0], L[:k], A[:k], N[:k], kid)
        F = n * (n + 1) // 2 * f_k(d) + n ** 3 / 3 + n ** 2 + 3 * n
        _lib.loglik_batch(X, Y[0], L, A, N, kid)     # first host call of a shape grows the library's workspaces
        t0 = time.perf_counter(); ll_h = _lib.loglik_batch(X, Y[0], L, A, N, kid); t_e2e = time.perf_counter() - t0
        out.append({"config": "C3", "kernel": kname, "n": n, "d": d, "S_per_gpu": S, "ms_per_step": ms,
                    "loglik_evals_per_s": S / (ms * 1e-3), "flop_per_eval": F,
                    "tflops": F * S / (ms * 1e-3) * 1e-12, "frac_of_dgemm_peak": F * S / (ms * 1e-3) * 1e-12 / peak(),
                    "e2e_host_call_evals_per_s": S / t_e2e, "parity_max_relerr_vs_oracle": relerr(ll[:k], ref),
                    "host_vs_dev_bitexact": bool(np.array_equal(ll_h, ll))})
    return out


def run_c5(torch, _lib, O, stream, args):
    from tests.util_problems import make_problem, relerr
    n, d, y_dim = 1024, 10, 4
    M = (1 << 22) if args.full else (1 << 19)
    X, Y, ls, amp, ns = make_problem(n, d, seed=1005, y_dim=y_dim)
    gps = [_lib.gp_fit(X, Y[i], ls[i], amp[i], ns[i], 2) for i in range(y_dim)]
    y_max = np.array([np.inf] + [float(np.quantile(Y[i], 0.7)) for i in range(1, y_dim)])
    coefs = np.array([1.0, 0, 0, 0])
    best = O.best_so_far(coefs, Y, y_max)
    gen = torch.Generator(device="cuda"); gen.manual_seed(5005)
    Xs = torch.rand((M, d), dtype=torch.float64, device="cuda", generator=gen)
    acq = torch.empty(M, dtype=torch.float64, device="cuda")
    res = {}

    def step():
        res["r"] = _lib.ei_score_dev(gps, y_dim, 1, Xs.data_ptr(), M, coefs, best, y_max, acq_ptr=acq.data_ptr())
    ms = timed(torch, stream, step, args.steps)
    k = 4096
    Xs_h = Xs[:k].cpu().numpy().T
    posts = [O.posterior_fit(X, Y[i], ls[i], amp[i], ns[i], 2) for i in range(y_dim)]
    ref, _, _ = O.ei_acquisition([posts], Xs_h, coefs, best, y_max)
    got = acq[:k].cpu().numpy()
    m = ref > 1e-30 * np.max(ref)                  # beyond that EI sits > 11 sigma in the tail: rel. error amplified by z^2
    a_all = acq.cpu().numpy()
    F = y_dim * (n * n + n * (3 * d + 12))
    out = {"config": "C5", "n": n, "d": d, "y_dim": y_dim, "M_per_gpu": M, "ms_per_step": ms,
           "candidates_per_s": M / (ms * 1e-3), "flop_per_candidate": F, "tflops": F * M / (ms * 1e-3) * 1e-12,
           "frac_of_dgemm_peak": F * M / (ms * 1e-3) * 1e-12 / peak(),
           "parity_max_relerr_vs_oracle": relerr(got[m], ref[m]), "parity_points": int(m.sum()),
           "argmax_matches_full_vector": bool(int(np.argmax(a_all)) == res["r"][1])}
    for g in gps:
        g.free()
    return [out]


def run_c4(torch, _lib, O, stream, args):
    from tests.util_problems import make_problem, relerr
    n, d = 4096, 4
    M = 8192 if args.full else 1024
    X, Y, ls, amp, ns = make_problem(n, d, seed=1004)
    theta = np.array([0.8, -0.3])
    mean_X = theta[0] * X[0] + theta[1]                       # Semiparametric: host-evaluated parametric mean
    y = Y[0] + mean_X
    gp = _lib.gp_fit(X, y - mean_X, ls[0], amp[0], ns[0], 2)
    best = float(np.max(y))
    rng = np.random.default_rng(4004)
    starts = (rng.permuted(np.tile(np.arange(M), (d, 1)), axis=1) + rng.random((d, M))) / M   # LHC in [0,1]^d
    Xs = torch.tensor(np.ascontiguousarray(starts.T), device="cuda")
    pm = torch.tensor(theta[0] * starts[0] + theta[1], device="cuda")          # m(x*) per start (y_dim = 1)
    pmg_h = np.zeros((M, d)); pmg_h[:, 0] = theta[0]                              # index (m*d + j)*y_dim + i
    pmg = torch.tensor(pmg_h, device="cuda")
    acq = torch.empty(M, dtype=torch.float64, device="cuda")
    grad = torch.empty((M, d), dtype=torch.float64, device="cuda")
    lb, ub = np.zeros(d), np.ones(d)

    def step_vg():
        _lib.ei_value_grad_dev([gp], 1, 1, Xs.data_ptr(), M, [1.0], best, None, acq.data_ptr(), grad.data_ptr(), lb=lb, ub=ub,
                               prior_mean_ptr=pm.data_ptr(), prior_mean_grad_ptr=pmg.data_ptr())

    def step_v():
        _lib.ei_score_dev([gp], 1, 1, Xs.data_ptr(), M, [1.0], best, None, lb=lb, ub=ub, acq_ptr=acq.data_ptr(),
                          prior_mean_ptr=pm.data_ptr())
    ms_vg = timed(torch, stream, step_vg, args.steps)
    g_dev = grad.cpu().numpy().T.copy(); a_dev = acq.cpu().numpy().copy()
    ms_v = timed(torch, stream, step_v, args.steps)
    k = 256
    post = O.posterior_fit(X, y - mean_X, ls[0], amp[0], ns[0], 2)
    a_ref, g_ref = O.ei_value_grad([post], starts[:, :k], [1.0], best, None,
                                   prior_mean_s=[theta[0] * starts[0, :k] + theta[1]],
                                   prior_mean_grad_s=[np.ascontiguousarray(pmg_h[:k].T)])
    m = a_ref > 1e-30 * np.max(a_ref)              # see run_c5: deeper in the tail EI's rel. error is z^2-amplified
    gscale = np.max(np.abs(g_ref[:, m]), axis=0)
    gerr = float(np.max(np.abs(g_dev[:, :k][:, m] - g_ref[:, m]) / gscale)) if m.any() else 0.0
    # the whole multi-start solve, device resident: 50 lock-step L-BFGS iterations over all starts
    aff = np.zeros((1, d + 1)); aff[0, 0] = theta[1]; aff[0, 1] = theta[0]
    _lib.ei_maximize_multistart([gp], 1, 1, starts, [1.0], best, None, lb, ub, iters=50, prior_mean_affine=aff)   # grows workspaces
    walls = []
    for _ in range(3):       # a single host-timed call is exposed to scheduling hiccups of the box: median of three
        t0 = time.perf_counter()
        Xo, fo, bx, bv, bi, evals = _lib.ei_maximize_multistart([gp], 1, 1, starts, [1.0], best, None, lb, ub, iters=50,
                                                                prior_mean_affine=aff)
        walls.append(time.perf_counter() - t0)
    t_opt = _max_over_ranks(torch, float(np.median(walls)) * 1e3) * 1e-3
    Fg = 2 * n * n + n * (9 * d + 16)
    Fv = n * n + n * (3 * d + 12)
    out = {"config": "C4", "n": n, "d": d, "starts_per_gpu": M, "ms_per_value_grad_iteration": ms_vg,
           "value_grad_evals_per_s": M / (ms_vg * 1e-3), "flop_per_start_iteration": Fg,
           "tflops_value_grad": Fg * M / (ms_vg * 1e-3) * 1e-12,
           "frac_of_dgemm_peak_value_grad": Fg * M / (ms_vg * 1e-3) * 1e-12 / peak(),
           "ms_per_value_only": ms_v, "frac_of_dgemm_peak_value_only": Fv * M / (ms_v * 1e-3) * 1e-12 / peak(),
           "lockstep_50_iterations_s": 50 * ms_vg * 1e-3,
           "on_device_multistart": {"iters": 50, "batched_value_grad_evals": int(evals), "wall_s": t_opt,
                                    "wall_s_of_3_calls": [round(w, 6) for w in walls],
                                    "best_value": float(bv), "start_value_max": float(np.max(a_dev)),
                                    "tflops": Fg * M * evals / t_opt * 1e-12,
                                    "frac_of_dgemm_peak": Fg * M * evals / t_opt * 1e-12 / peak()},
           "on_device_multistart_frac": Fg * M * evals / t_opt * 1e-12 / peak(),
           "parity_value_max_relerr": relerr(a_dev[:k][m], a_ref[m]), "parity_grad_max_err_rel_to_grad_norm": gerr,
           "parity_points": int(m.sum())}
    gp.free()
    return [out]


def run_all(torch, _lib, stream, steps=3, full=False, configs=("c3", "c4", "c5"), dist=None, world=1, peak_tflops=None):
    """In-process entry for bench.py: list of per-config result dicts (per-GPU shares; with `dist` every rank runs its
    share and each time is the max over ranks, so the whole-job rate is the per-GPU rate x world)."""
    from oracle import boss_oracle as O
    args = argparse.Namespace(steps=steps, full=full)
    _DIST[0] = dist
    _PEAK[0] = peak_tflops
    out = []
    try:
        for c in configs:
            for line in {"c3": run_c3, "c4": run_c4, "c5": run_c5}[c](torch, _lib, O, stream, args):
                line["full_job_on_one_gpu"] = bool(full)
                line["n_gpus"] = world
                out.append(line)
    finally:
        _DIST[0] = None
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c3,c4,c5")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    import torch
    import boss_b200  # noqa: F401
    from boss_b200 import _lib
    from oracle import boss_oracle as O
    torch.cuda.set_device(0)
    _lib.init(0)
    stream = torch.cuda.ExternalStream(_lib.stream_ptr(), device=torch.device("cuda", 0))
    for c in args.configs.split(","):
        for line in {"c3": run_c3, "c4": run_c4, "c5": run_c5}[c](torch, _lib, O, stream, args):
            line["full_job_on_one_gpu"] = bool(args.full)
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
