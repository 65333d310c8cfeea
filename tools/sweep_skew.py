"""Warp-group skew of the 8192- and 16384-point overlap-save kernels (llz_cuda_tune "fft8k_skew" / "fft16k_skew": cycles by
which warps 4..7 of a CTA trail warps 0..3 after the first exchange).  Results: profiles/r01_sweep_skew.txt.
Run under gpurun: python tools/sweep_skew.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llzlab_b200 as z  # noqa: E402

C_, n = 16, 16_000_000
for dtype, tdt in ((z.F64, torch.float64), (z.F32, torch.float32)):
    dx = torch.randn(C_, n, dtype=tdt, device="cuda")
    dy = torch.empty_like(dx)
    for size, taps, var in (("8192", 4095, "fft8k_skew"), ("8192", 2049, "fft8k_skew"),
                            ("16384", 4095, "fft16k_skew"), ("16384", 8191, "fft16k_skew")):
        row = [f"{'f64' if dtype == z.F64 else 'f32'} {size:>5s} taps {taps:5d}:"]
        for skew in (0, 600, 900, 1100, 1300, 1500, 1800, 2200):
            z.tune(var, skew)
            bank = z.FirBank(C_, dtype, kind=z.LPF, flt_len=taps, fc1=0.2, algo=z.FIR_FFT)
            bank.set_fft_size(int(size))
            for _ in range(3):
                bank.run(dx, n, dy, n, n)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                bank.run(dx, n, dy, n, n)
            e1.record()
            torch.cuda.synchronize()
            row.append(f"{skew}:{C_ * n * 5 / e0.elapsed_time(e1) / 1e6:6.1f}")
            bank.close()
        z.tune(var, -1)
        print("  ".join(row), flush=True)
