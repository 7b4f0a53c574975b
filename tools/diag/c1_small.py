"""C1-style fitter batch: S hyper-parameter vectors x loglik(n=30, d=2) -- warp-register path timing."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import boss_b200
from boss_b200 import _lib
from tests.util_problems import make_problem, make_hyper_samples
_lib.init(0)
st = torch.cuda.ExternalStream(_lib.stream_ptr(), device=torch.device("cuda", 0))
for n, d, S in ((30, 2, 4096), (30, 2, 65536), (32, 8, 65536), (20, 2, 200)):
    X, Y, _, _, _ = make_problem(n, d, seed=1)
    L, A, N = make_hyper_samples(S, d, seed=2)
    tX = torch.tensor(np.ascontiguousarray(X.T), device="cuda"); ty = torch.tensor(Y[0], device="cuda")
    tL = torch.tensor(L, device="cuda"); tA = torch.tensor(A, device="cuda"); tN = torch.tensor(N, device="cuda")
    out = torch.empty(S, dtype=torch.float64, device="cuda")
    f = lambda: _lib.loglik_batch_dev(tX.data_ptr(), d, n, ty.data_ptr(), 0, tL.data_ptr(), tA.data_ptr(), tN.data_ptr(), 0, S, out.data_ptr())
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): f()
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    t0 = time.perf_counter(); _lib.loglik_batch(X, Y[0], L, A, N, 0); th = time.perf_counter() - t0
    print(f"n={n} d={d} S={S}: {ms*1e3:.1f} us/batch dev-resident = {S/ms*1e3:.3e} evals/s ; host-call {th*1e6:.0f} us")
