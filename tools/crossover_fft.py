"""Where does the 8192-point overlap-save kernel overtake the 1024-point one?  (tuning of kFirFft8kMinTapsAuto)
Run under gpurun: python tools/crossover_fft.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llzlab_b200 as z  # noqa: E402

C_, n = 64, 4_000_000
for dtype, tdt in ((z.F64, torch.float64), (z.F32, torch.float32)):
    dx = torch.randn(C_, n, dtype=tdt, device="cuda")
    dy = torch.empty_like(dx)
    for taps in (129, 257, 385, 449, 513, 577, 641, 769, 897):
        row = [f"{'f64' if dtype == z.F64 else 'f32'} taps {taps:4d}:"]
        for size in ("1024", "8192", "direct"):
            if size == "direct":
                bank = z.FirBank(C_, dtype, kind=z.LPF, flt_len=taps, fc1=0.2, algo=z.FIR_DIRECT)
            else:
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
            row.append(f"{size} {C_ * n * 5 / e0.elapsed_time(e1) / 1e6:8.1f} Gs/s")
            bank.close()
        print("  ".join(row))
