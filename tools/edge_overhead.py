"""Fixed cost per call of the overlap-save FIR path (edge items + history): time C2's bank at several stream lengths
and fit t = a + b*n.  Run under gpurun: python tools/edge_overhead.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llzlab_b200 as z  # noqa: E402

C_ = 1024
for dtype, tdt in ((z.F64, torch.float64), (z.F32, torch.float32)):
    bank = z.FirBank(C_, dtype, kind=z.LPF, flt_len=127, fc1=0.23, win=z.HAMMING)
    rows = []
    for n in (120_000, 240_000, 480_000, 960_000):
        dx = torch.randn(C_, n, dtype=tdt, device="cuda")
        dy = torch.empty_like(dx)
        for _ in range(3):
            bank.run(dx, n, dy, n, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            bank.run(dx, n, dy, n, n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        rows.append((n, ms))
        print(f"{'f64' if dtype == z.F64 else 'f32'} n={n:7d}: {ms:7.4f} ms  {C_ * n / ms / 1e6:7.1f} Gsamples/s", flush=True)
        del dx, dy
    (n0, t0), (n1, t1) = rows[1], rows[3]
    b = (t1 - t0) / (n1 - n0)
    print(f"   fit: fixed {t0 - b * n0:6.4f} ms per call, {1e-6 * C_ / b:7.1f} Gsamples/s asymptotic")
    bank.close()
