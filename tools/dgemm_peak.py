"""cuBLAS DGEMM throughput on this GPU (the FP64 roofline denominator MEASURED_PEAKS.json lacks)."""
import json, sys, time
import torch

def run(n=8192, burst=10, sustain_s=3.0):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(burst):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fl = 2.0 * n ** 3
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.time(); k = 0
    e0.record()
    while time.time() - t0 < sustain_s:
        for _ in range(4):
            torch.matmul(a, b, out=c); k += 1
        torch.cuda.synchronize()
    e1.record(); e1.synchronize()
    sus = e0.elapsed_time(e1) / k
    return {"n": n, "dgemm_tflops_burst": fl / best * 1e-9, "dgemm_tflops_sustained": fl / sus * 1e-9}

if __name__ == "__main__":
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    for n in (4096, 8192):
        out[f"n{n}"] = run(n)
    print(json.dumps(out))
