#!/usr/bin/env python
"""Small-batch scoring latency: wall time of one value (+ gradient) call for batch sizes 1 .. 1024 through the wide
(128-candidate), narrow (32-candidate) and quarter-row-block kernels (BOSS_SCORE_PATH), and a bitwise comparison of
the three.  Calibrates the dispatch cost model in boss_b200.cu (NQ_COST_*).  Prints one JSON line per (n, batch)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import boss_b200  # noqa: F401
    from boss_b200 import _lib
    from tests.util_problems import make_problem
    _lib.init(0)
    sizes = [int(x) for x in os.environ.get("SB_SIZES", "1,8,32,64,128,256,512,1024").split(",")]
    shapes = [tuple(int(v) for v in t.split("x")) for t in os.environ.get("SB_SHAPES", "2048x8,4096x4").split(",")]
    for n, d in shapes:
        X, Y, ls, amp, ns = make_problem(n, d, seed=77)
        gp = _lib.gp_fit(X, Y[0], ls[0], amp[0], ns[0], 2)
        best = float(np.max(Y[0]))
        lb, ub = np.zeros(d), np.ones(d)
        rng = np.random.default_rng(5)
        for M in sizes:
            Xs = torch.tensor(rng.random((M, d)), device="cuda")
            acq = torch.empty(M, dtype=torch.float64, device="cuda")
            grad = torch.empty((M, d), dtype=torch.float64, device="cuda")
            row = {"n": n, "d": d, "batch": M}
            ref = {}
            for path in ("wide", "narrow", "quarter", "auto"):
                if path == "auto":
                    os.environ.pop("BOSS_SCORE_PATH", None)
                else:
                    os.environ["BOSS_SCORE_PATH"] = path

                def vg():
                    _lib.ei_value_grad_dev([gp], 1, 1, Xs.data_ptr(), M, [1.0], best, None, acq.data_ptr(), grad.data_ptr(), lb=lb, ub=ub)

                def v():
                    _lib.ei_score_dev([gp], 1, 1, Xs.data_ptr(), M, [1.0], best, None, lb=lb, ub=ub, acq_ptr=acq.data_ptr())
                for name, fn in (("v", v), ("vg", vg)):
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    reps = 20
                    t0 = time.perf_counter()
                    for _ in range(reps):
                        fn()
                    torch.cuda.synchronize()
                    row[f"{path}_{name}_us"] = round((time.perf_counter() - t0) / reps * 1e6, 1)
                    out = (acq.cpu().numpy().copy(), grad.cpu().numpy().copy() if name == "vg" else None)
                    if path == "wide":
                        ref[name] = out
                    else:
                        same = np.array_equal(out[0], ref[name][0]) and (out[1] is None or np.array_equal(out[1], ref[name][1]))
                        row[f"{path}_{name}_bits_equal_wide"] = bool(same)
            # where the time of the (auto-dispatched) value + gradient call goes: CUDA-event time per kernel class
            _lib.set_timing(True)
            vg()
            row["auto_vg_kernel_us"] = {k: (round(_lib.last_kernel_ms(w)[0] * 1e3, 1), _lib.last_kernel_ms(w)[1])
                                        for k, w in (("trmm_class", 0), ("xcov_grad_class", 1), ("whole_call", 3))}
            _lib.set_timing(False)
            print(json.dumps(row), flush=True)
        gp.free()
    os.environ.pop("BOSS_SCORE_PATH", None)


if __name__ == "__main__":
    main()
