#!/usr/bin/env python
"""Single-process multi-GPU through the C ABI (boss_init_multi): the layout a Julia caller has.

One process drives N GPUs; the library replicates the fit, deals candidate blocks / hyper-parameter samples / starts
to the devices (one host thread + stream set per device) and reduces the (best value, index) pairs on the host.
Workloads = bench.py's: C2 per-GPU share x N (n = 2048, d = 8, 2 Mi candidates per GPU, host arrays) and the headline
log-likelihood batch (256 samples per GPU).  Prints one JSON line with the N = 1 and N = all results and the scaling.

    python tools/bench_multi_inlib.py [--gpus N] [--steps 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import boss_b200  # noqa: F401
    from boss_b200 import _lib
    from tests.util_problems import make_hyper_samples, make_problem
    N = args.gpus or torch.cuda.device_count()
    n, d, kid = 2048, 8, 2
    M1, S1 = 1 << 21, 256
    X, Y, ls, amp, ns = make_problem(n, d, seed=1002)
    y = Y[0]
    best = float(np.max(y))
    out = {"what": "single-process multi-GPU through boss_init_multi (host arrays, pageable)", "n_gpus": N}
    # one sample set for every device count: the N = 1 call evaluates the first 256 samples of the N-device call
    La, Aa, Na = make_hyper_samples(S1 * N, d, seed=3003)
    La = np.ascontiguousarray(La)
    for nd in ([1, N] if N > 1 else [1]):
        if nd == 1:
            _lib.init(0)
        else:
            _lib.init_multi(nd)
        gp = _lib.gp_fit(X, y, ls[0], float(amp[0]), float(ns[0]), kid)
        M, S = M1 * nd, S1 * nd
        Xs = np.random.default_rng(2002).random((M, d)).T          # d x M view of an M x d pageable array
        L, A, Nn = np.ascontiguousarray(La[:S]), Aa[:S].copy(), Na[:S].copy()      # ls is S x d (sample-major)
        res = {}
        for name, fn in (("score", lambda: _lib.ei_score([gp], 1, 1, Xs, [1.0], best, None, want_acq=False)),
                         ("loglik", lambda: _lib.loglik_batch(X, y, L, A, Nn, kid))):
            for _ in range(2):
                r = fn()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                r = fn()
            dt = (time.perf_counter() - t0) / args.steps
            res[name] = {"ms_per_call": dt * 1e3, "per_s": (M if name == "score" else S) / dt}
            if name == "score":
                res[name]["argmax"] = [r[1], int(r[2])]
            else:
                res[name]["checksum"] = float(np.sum(r[:S1]))
        out[f"n{nd}"] = res
        gp.free()
    if N > 1:
        out["scaling"] = {k: out[f"n{N}"][k]["per_s"] / out["n1"][k]["per_s"] / N for k in ("score", "loglik")}
        out["loglik_first_shard_bit_identical"] = out["n1"]["loglik"]["checksum"] == out[f"n{N}"]["loglik"]["checksum"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
